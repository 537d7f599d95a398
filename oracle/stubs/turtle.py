"""Stand-in for the stdlib `turtle` module (needs tkinter, absent here); vision_transformer.py:10 imports an unused name."""


def back(*a, **k):
    raise RuntimeError("turtle stub")
