"""No-op stand-in for ml_logger (absent from this image); test infrastructure only."""


class _Ctx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _Logger:
    prefix = "shim"

    def __getattr__(self, name):
        def _noop(*a, **k):
            return _Ctx()
        return _noop


logger = _Logger()
