"""Import the pure-numpy modules of the reference (predef_path, transformations,
obstacles, test_scenarios) in a container that has neither pymunk, pygame nor
matplotlib, by stubbing those packages.  ONLY used by make_golden.py (run once,
in the authoring container, where /root/reference exists).  Nothing under
tests/ that runs on the GPU box imports this file.
"""
import sys
import types

REF_DIR = "/root/reference/drone_2d_custom_gym_env"


class _Anything:
    """Constructor-only stand-in: accepts any call / attribute."""

    def __init__(self, *a, **k):
        self._a = a
        self._k = k

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Anything()

    def __call__(self, *a, **k):
        return _Anything(*a, **k)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    m.__dict__.update(attrs)

    def _module_getattr(n):
        if n.startswith("__"):
            raise AttributeError(n)
        return _Anything()

    m.__getattr__ = _module_getattr  # type: ignore
    sys.modules[name] = m
    return m


def install_stubs():
    mpl = _stub("matplotlib")
    mpl.pyplot = _stub("matplotlib.pyplot")
    mpl.patches = _stub("matplotlib.patches")

    class Body(_Anything):
        STATIC = 2
        DYNAMIC = 0

        def __init__(self, *a, **k):
            self.position = (0.0, 0.0)

    class Circle(_Anything):
        def __init__(self, body, radius, *a, **k):
            self.body = body
            self.radius = radius

    class Space(_Anything):
        def add(self, *a):
            pass

    pm = _stub("pymunk", Body=Body, Circle=Circle, Space=Space, Vec2d=tuple)
    pm.pygame_util = _stub("pymunk.pygame_util")
    _stub("pygame", Color=lambda c: c)


def load():
    install_stubs()
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)
    import predef_path, transformations, obstacles, test_scenarios  # noqa
    return predef_path, transformations, obstacles, test_scenarios
