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


# ---------------------------------------------------------------- image frames
#: thing colours as the reference's constructors pass them (things.py:10,34,46,57-60; players/*.py create())
THING_COLORS = {"box": "yellow", "wall": "white", "zombie": "green", "objective": "blue", "agent": "blue",
                "terminator": "cyan", "sniper": "yellow", "troll": "blue", "hamster": "yellow", "randoman": "red"}
#: the device keeps WHERE dead bodies lie, not whose they are (the reference paints a body in its owner's colour:
#: 'zombie remains' green, a dead player in the player's colour): bodies are painted as zombie remains
DEAD_BODY_COLOR = "green"


class ImageRenderer(object):
    """An RGB frame of one world of the batch with the geometry and the shapes of the reference's OpenCV renderer
    (zombsole/renderer.py:97-277, ``OpencvRenderer.render``): 10 x 10 pixel cells; walls filled, boxes outlined and
    crossed, zombies and players solid ellipses, objectives outlined, dead bodies an ellipse outline with a cross; below
    the map the counters line and one life bar per player.  The player lines' text carries each player's ``status``,
    which the device does not keep: the text is drawn without it.  A debugging aid built from the device state through
    the object views of things.py; not on the step path.  Needs Pillow (as the reference's renderer does)."""

    def __init__(self, gridwidth, gridheight, cellwidth=10, cellheight=10):
        self.cw, self.ch = cellwidth, cellheight
        self.imagewidth, self.imageheight = cellwidth * gridwidth, cellheight * gridheight
        self.lifebar_width = 20

    def _cell(self, x, y, w=1, h=1):
        return [(x * self.cw, y * self.ch), ((x + w) * self.cw, (y + h) * self.ch)]

    def _cross(self, img, x, y, color, w=1, h=1, width=1):
        (x0, y0), (x1, y1) = self._cell(x, y, w, h)
        img.line([(x0, y0), (x1, y1)], fill=color, width=width)
        img.line([(x1, y0), (x0, y1)], fill=color, width=width)

    def _thing(self, img, x, y, kind, color):
        if kind == "wall":
            img.rectangle(self._cell(x, y), fill=color, outline=None)
        elif kind == "box":
            img.rectangle(self._cell(x, y), fill=None, outline=color, width=2)
            self._cross(img, x, y, color)
        elif kind == "objective":
            img.rectangle(self._cell(x, y), fill=None, outline=color, width=2)
        elif kind == "dead body":
            img.ellipse(self._cell(x, y), fill=None, outline=color, width=1)
            self._cross(img, x, y, color)
        else:  # zombies and players
            img.ellipse(self._cell(x, y), fill=color, width=1)

    def _lifebar(self, img, x, y, player, color):
        pixels = int(self.lifebar_width * self.cw * player.life / player.MAX_LIFE)
        img.rectangle(self._cell(x, y, self.lifebar_width, 1), fill=None, outline=color if pixels > 0 else "red", width=2)
        if pixels > 0:
            img.rectangle([(x * self.cw, y * self.ch), (x * self.cw + pixels, (y + 1) * self.ch)], fill=color, outline=None, width=2)
        else:
            self._cross(img, x, y, "red", self.lifebar_width, 1, width=3)

    def draw_image(self, game, with_text=True):
        """-> uint8 array [height, width, 3] (RGB)."""
        import numpy as np
        from PIL import Image, ImageDraw
        world = game.world
        things, deco = world.things, world.decoration
        width, height = world.size
        image = Image.new("RGB", (self.imagewidth, self.imageheight), "black")
        img = ImageDraw.Draw(image)
        for x in range(width):  # (x-major like the reference: neighbouring shapes share their border pixels)
            for y in range(height):
                t = things.get((x, y))
                if t is not None:
                    self._thing(img, x, y, t.name if t.name in ("wall", "box", "zombie") else "player", THING_COLORS.get(t.name, "red"))
                elif (x, y) in deco:
                    kind = deco[(x, y)]
                    self._thing(img, x, y, kind, THING_COLORS["objective"] if kind == "objective" else DEAD_BODY_COLOR)
        if with_text:
            img.text((0, height * self.ch), "ticks: %d deaths: %d" % (world.t, world.deaths), font=None, fill="yellow", anchor="la", font_size=14)
        players = sorted(game.agents, key=lambda a: a.agent_id) + sorted(game.players, key=lambda p: p.name)
        for idx, player in enumerate(players):
            color = THING_COLORS.get(player.name, "red")
            self._lifebar(img, 1, height + 2 + idx, player, color)
            if with_text:
                stats = u"%s <%i %s %s>: -" % (player.name, player.life, str(player.position), player.weapon.name)
                img.text(((1 + self.lifebar_width + 1) * self.cw, (height + 2 + idx) * self.ch), stats, font=None, fill=color, anchor="la", font_size=8)
        return np.array(image)
