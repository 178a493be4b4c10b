"""Host-side views of device state with the reference's object protocol.

The reference's tests and callers reach into ``env.game.world.things``, ``env.game.agents[i].life``
and friends (tests/test_game.py:23-38,54-66; tests/test_multiagent_env.py:108-121).  These classes
give the same names over the SoA state in HBM: reading a property copies the value from the
device, assigning ``life`` writes it back.  They are debugging / compatibility sugar, not the
hot path — the hot path never leaves the device.
"""
from . import abi


class Thing(object):
    MAX_LIFE = 1
    name = "thing"
    icon_basic = "?"
    ask_for_actions = False
    is_decoration = False

    def __init__(self, engine, env, index):
        self._eng, self._env, self._index = engine, env, index

    def __repr__(self):
        return "<%s %s life=%s>" % (type(self).__name__, self.position, self.life)

    def __eq__(self, other):
        return type(self) is type(other) and (self._env, self._index) == (other._env, other._index)

    def __hash__(self):
        return hash((type(self).__name__, self._env, self._index))


class _Static(Thing):
    @property
    def position(self):
        x, y, _ = self._eng.map.statics[self._index]
        return (x, y)

    @property
    def life(self):
        return int(self._eng.fields["static_life"][self._env, self._index].item())

    @life.setter
    def life(self, value):
        self._eng.fields["static_life"][self._env, self._index] = int(value)
        self._eng.state_written()


class Box(_Static):
    MAX_LIFE = 10
    name = "box"
    icon_basic = "@"


class Wall(_Static):
    MAX_LIFE = 200
    name = "wall"
    icon_basic = "#"


class Weapon(object):
    _BY_CODE = {1: ("ZombieClaws", 1.5, (5, 10)), 10: ("Knife", 1.5, (5, 10)), 11: ("Axe", 1.5, (75, 100)),
                12: ("Gun", 6, (10, 50)), 13: ("Rifle", 10, (25, 75)), 14: ("Shotgun", 3, (75, 100))}

    def __init__(self, code):
        self.name, self.max_range, self.damage_range = self._BY_CODE[int(code)]


class FightingThing(Thing):
    MAX_LIFE = 100
    ask_for_actions = True

    @property
    def position(self):
        f = self._eng.fields
        return (int(f["x"][self._env, self._index].item()), int(f["y"][self._env, self._index].item()))

    @property
    def life(self):
        return int(self._eng.fields["life"][self._env, self._index].item())

    @life.setter
    def life(self, value):
        self._eng.fields["life"][self._env, self._index] = int(value)
        self._eng.state_written()

    @property
    def in_world(self):
        return bool(int(self._eng.fields["meta"][self._env, self._index].item()) & 0x80)

    @property
    def weapon(self):
        return Weapon(int(self._eng.fields["meta"][self._env, self._index].item()) & 15)


class Zombie(FightingThing):
    name = "zombie"
    icon_basic = "x"


class Player(FightingThing):
    name = "player"
    icon_basic = "P"


class Terminator(Player):
    name = "terminator"


class Agent(Player):
    name = "agent"
    icon_basic = "A"

    def __init__(self, engine, env, index, agent_id):
        super().__init__(engine, env, index)
        self.agent_id = agent_id
        self.thing_type = "agent"


class World(object):
    """Read-only view with the reference's World attributes (zombsole/core.py:12-20)."""

    def __init__(self, game):
        self._game = game

    @property
    def size(self):
        return self._game.map.size

    def _scalar(self, idx):
        return int(self._game._eng.fields["scalars"][self._game._env, idx].item())

    @property
    def t(self):
        return self._scalar(abi.S_T)

    @property
    def deaths(self):
        return self._scalar(abi.S_DEATHS)

    @property
    def zombie_deaths(self):
        return self._scalar(abi.S_ZOMBIE_DEATHS)

    @property
    def things(self):
        """{(x, y): thing} in World.things order (statics first, then mobiles by dict-order stamp)."""
        eng, env = self._game._eng, self._game._env
        f = {k: v[env].cpu().numpy() for k, v in eng.fields.items()}
        fresh = bool(f["scalars"][abi.S_FLAGS] & 1)
        out = {}
        for i, (x, y, label) in enumerate(eng.map.statics):
            if f["static_life"][i] > 0 or fresh:
                out[(x, y)] = (Box if label == 1 else Wall)(eng, env, i)
        live = [s for s in range(eng.M) if f["meta"][s] & 0x80]
        live.sort(key=lambda s: f["stamp"][s])
        for s in live:
            out[(int(f["x"][s]), int(f["y"][s]))] = self._game._slot_thing(s)
        return out

    @property
    def decoration(self):
        eng, env = self._game._eng, self._game._env
        bits = eng.fields["dead_body"][env].cpu().numpy().view("uint8")
        w = eng.map.size[0]
        out = {tuple(o): "objective" for o in eng.map.objectives}
        for c in range(eng.cells):
            if (bits[c >> 3] >> (c & 7)) & 1:
                out[(c % w, c // w)] = "dead body"
        return out

    def within_bounds(self, position):
        return 0 <= position[0] < self.size[0] and 0 <= position[1] < self.size[1]


class Game(object):
    """View with the reference's Game attributes (zombsole/game.py:109-201) for one env of a batch."""

    def __init__(self, engine, env, rules_name, player_names, agent_ids, initial_zombies, minimum_zombies):
        self._eng, self._env = engine, env
        self.rules_name = rules_name
        self.player_names = list(player_names)
        self.agent_ids = list(agent_ids)
        self.initial_zombies, self.minimum_zombies = initial_zombies, minimum_zombies
        self.map = engine.map
        self.world = World(self)
        self.players = [Terminator(engine, env, i) for i in range(len(player_names))]
        for pl, name in zip(self.players, self.player_names):
            pl.name = name  # every scripted player is created with its module's name (players/*.py: create())
        self.agents = [Agent(engine, env, len(player_names) + i, aid) for i, aid in enumerate(agent_ids)]

    def _slot_thing(self, s):
        np_ = len(self.players)
        if s < np_:
            return self.players[s]
        if s < np_ + len(self.agents):
            return self.agents[s - np_]
        return Zombie(self._eng, self._env, s)

    def get_all_players(self):
        return self.players + self.agents

    def draw_text(self, use_basic_icons=True):
        """The frame Game.draw() would print (game.py:236-238, renderer.py:45-88), as a string."""
        from .renderer import TerminalRenderer
        return TerminalRenderer(use_basic_icons).draw_text(self)

    def draw_image(self, with_text=True):
        """The frame the reference's OpenCV renderer would show (renderer.py:241-277), as an RGB uint8 array."""
        from .renderer import ImageRenderer
        w, h = self.map.size
        return ImageRenderer(w, h + 2 + len(self.players) + len(self.agents)).draw_image(self, with_text)

    def draw(self):
        print(self.draw_text())

    def get_agents_health(self):
        return sum(t.life for t in self.agents)

    def get_players_health(self):
        return sum(t.life for t in self.players)
