"""Stand-in for termcolor (see ../README.md): colouring is a no-op."""


def colored(text, *args, **kwargs):
    return text


def cprint(text, *args, **kwargs):
    print(text)
