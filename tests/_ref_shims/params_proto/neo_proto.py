"""Stand-in for params_proto.neo_proto: class-attribute namespaces.

Behaviour relied on by the reference (SURVEY.md section 5 "Config / flags"):
  * `class X(PrefixProto, cli=False)` / `class Y(ParamsProto, cli=False, prefix="...")`
  * `vars(X)` returns a FRESH deep dict of the public attributes
    (legged_robot.py:1420-1421 mutates the result).
  * `X._update(dict)` bulk assignment (scripts/play.py:34-46).
"""
import copy
import inspect


_REAL_DICT = type.__dict__["__dict__"]


class Meta(type):
    def __new__(mcs, name, bases, ns, cli=True, prefix=None, **kw):
        return super().__new__(mcs, name, bases, ns)

    def __init__(cls, name, bases, ns, cli=True, prefix=None, **kw):
        super().__init__(name, bases, ns)

    @property
    def __dict__(cls):
        out = {}
        for klass in reversed(cls.__mro__):
            if klass in (object,) or klass.__name__ in ("PrefixProto", "ParamsProto"):
                continue
            for k, v in _REAL_DICT.__get__(klass).items():
                if k.startswith("_") or inspect.isfunction(v) or isinstance(v, (classmethod, staticmethod, property)):
                    continue
                out[k] = v if isinstance(v, type) else copy.deepcopy(v)
        return out

    def _update(cls, d=None, **kw):
        d = dict(d or {}, **kw)
        for k, v in d.items():
            if isinstance(v, dict) and isinstance(getattr(cls, k, None), Meta):
                getattr(cls, k)._update(v)
            else:
                setattr(cls, k, v)


class PrefixProto(metaclass=Meta):
    pass


class ParamsProto(metaclass=Meta):
    pass
