class App:
    def default(self, *a, **k):
        return (lambda f: f) if not (a and callable(a[0])) else a[0]
    def command(self, *a, **k):
        return (lambda f: f) if not (a and callable(a[0])) else a[0]
    def __call__(self, *a, **k):
        raise SystemExit("cyclopts shim: CLI not available")
class Parameter:
    def __init__(self, *a, **k): pass
class _V:
    def Path(self, *a, **k): return None
validators = _V()
