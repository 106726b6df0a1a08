"""Import-time only (atc_gym.py:7, rendering.py:2 of the reference); render() is never called."""


class Geom(object):
    def __init__(self):
        self.attrs = []

    def render(self):
        pass
