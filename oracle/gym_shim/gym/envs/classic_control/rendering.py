"""Dummy of gym's pyglet viewer: the reference imports it at module top; nothing here draws."""


class _Geom(object):
    def __init__(self, *a, **k):
        self.color = (0, 0, 0)

    def set_color(self, r, g, b):
        self.color = (r, g, b)


class FilledPolygon(_Geom):
    pass


class Viewer(object):
    def __init__(self, width, height, display=None):
        self.width, self.height, self.geoms = width, height, []

    def add_geom(self, geom):
        self.geoms.append(geom)

    def render(self, return_rgb_array=False):
        return None

    def close(self):
        pass


class SimpleImageViewer(object):
    def __init__(self, display=None):
        self.isopen = False

    def imshow(self, arr):
        self.isopen = True

    def close(self):
        self.isopen = False
