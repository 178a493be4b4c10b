"""Text frames of one world of the batch, laid out like the reference's terminal renderer
(zombsole/renderer.py:24-94, ``TerminalRenderer._draw``): the map with one icon per cell (things over decorations),
the counters line, and one line per player with the reference's life bar.  A debugging aid — a visual diff of a
single env when a parity test fails — built from the device state through the object views of things.py; it is
not on the step path.  What the device does not keep is each player's ``status`` text (set inside the reference's
``next_step`` implementations): the status column shows ``-`` as the reference does for an empty status.
"""
try:  # pragma: no cover - depends on the image
    from termcolor import colored
except ImportError:
    def colored(text, *args, **kwargs):
        return text

ICONS = {"box": u"☒", "wall": u"▓", "zombie": u"⨰", "player": u"⨰", "dead body": u"☠",
         "objective": u"░"}
ICONS_BASIC = {"box": u"@", "wall": u"#", "zombie": u"x", "player": u"P", "agent": u"A", "dead body": u"=", "objective": u"*"}
COLORS = {"box": "yellow", "wall": "white", "zombie": "green", "dead body": "white", "objective": "blue"}


class TerminalRenderer(object):
    def __init__(self, use_basic_icons=True, debug=False):
        self.use_basic_icons = use_basic_icons
        self.debug = debug

    def _icon(self, kind):
        if self.use_basic_icons:
            return ICONS_BASIC.get(kind, u"P")
        return ICONS.get(kind, ICONS["player"])

    def draw_text(self, game, color=False):
        """The frame as a string (renderer.py:45-88)."""
        world = game.world
        things, deco = world.things, world.decoration
        width, height = world.size
        paint = colored if color else (lambda text, *a, **k: text)
        rows = []
        for y in range(height):
            row = []
            for x in range(width):
                t = things.get((x, y))
                if t is not None:
                    kind = "agent" if getattr(t, "thing_type", None) == "agent" else \
                        ("player" if t.icon_basic == "P" else t.name)
                    row.append(paint(self._icon(kind), COLORS.get(t.name, "red")))
                elif (x, y) in deco:
                    row.append(paint(self._icon(deco[(x, y)]), COLORS[deco[(x, y)]]))
                else:
                    row.append(u" ")
            rows.append(u"".join(row))
        screen = u"\n".join(rows)
        screen += u"\nticks: %i deaths: %i, zombie deaths: %i" % (world.t, world.deaths, world.zombie_deaths)
        # Game.draw (game.py:236-238): agents by agent_id, then the scripted players by name
        players = sorted(game.agents, key=lambda a: a.agent_id) + sorted(game.players, key=lambda p: p.name)
        for player in sorted(players, key=lambda p: p.name):
            life = player.life
            if life > 0:
                n = int((10.0 / player.MAX_LIFE) * life)
                life_bar = u"♥ %s%s" % (n * u"█", (10 - n) * u"░")
            else:
                life_bar = u"☠ [dead]"
            stats = u"%s %s <%i %s %s>: %s" % (life_bar, player.name, life, str(player.position), player.weapon.name, u"-")
            screen += u"\n" + paint(stats, "red")
        return screen

    def render(self, game):
        print(self.draw_text(game, color=True))
