"""Stub of the ``visdom`` client: jTransUP/utils/visuliazer.py:1 imports it unconditionally.
Runs always pass ``-nohas_visualization``; constructing it is an error on purpose."""


class Visdom(object):
    def __init__(self, *a, **kw):
        raise RuntimeError("visdom is not available in this image; pass -nohas_visualization")
