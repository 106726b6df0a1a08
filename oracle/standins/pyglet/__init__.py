"""Empty stand-in: the reference imports pyglet at module import only (rendering.py:1)."""
