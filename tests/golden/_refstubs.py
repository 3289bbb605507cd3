"""Stub modules so that `import astro` from /root/reference works in the build
container (flask, lru-dict and tensorboardX are not installed; none of them is on
the core.step / get_features path).  Only used by make_golden.py, never at test time.
"""
import sys
import types


def install():
    flask = types.ModuleType('flask')

    class Flask:
        def __init__(self, *a, **k):
            pass

        def route(self, *a, **k):
            return lambda f: f
    flask.Flask = Flask
    flask.request = None
    flask.jsonify = lambda *a, **k: None
    flask.render_template = lambda *a, **k: None
    lru = types.ModuleType('lru')
    lru.LRU = lambda n: {}
    tbx = types.ModuleType('tensorboardX')

    class SummaryWriter:
        def __init__(self, *a, **k):
            pass

        def add_scalar(self, *a, **k):
            pass
    tbx.SummaryWriter = SummaryWriter
    for name, mod in (('flask', flask), ('lru', lru), ('tensorboardX', tbx)):
        sys.modules.setdefault(name, mod)
    if '/root/reference' not in sys.path:
        sys.path.insert(0, '/root/reference')
