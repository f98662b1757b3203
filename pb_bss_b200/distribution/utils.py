"""Model containers (pb_bss/distribution/utils.py:118-190 of the reference)."""
import difflib


class _ProbabilisticModel:
    """Dataclass mix-in: ``to_dict`` / ``from_dict`` round trip and helpful
    AttributeErrors, like the reference's base class of the same name."""

    def to_dict(self):
        out = {}
        for k in self.__dataclass_fields__.keys():
            v = getattr(self, k)
            out[k] = v.to_dict() if isinstance(v, _ProbabilisticModel) else v
        return out

    @classmethod
    def from_dict(cls, d):
        assert cls.__dataclass_fields__.keys() == d.keys(), (
            cls.__dataclass_fields__.keys(), d.keys())
        return cls(**d)

    def __getattr__(self, name):
        fields = list(self.__dataclass_fields__.keys())
        similar = difflib.get_close_matches(name, fields) or fields
        raise AttributeError(
            f'{self.__class__.__name__!r} object has no attribute {name!r}.\n'
            f'Close matches: {similar}')
