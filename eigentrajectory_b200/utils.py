"""Small host-side helpers mirrored from the reference's ``utils/utils.py``."""


class DotDict(dict):
    """Dictionary with attribute access; a missing key reads as ``None`` (utils/utils.py:32-39)."""

    __getattr__ = dict.get
    __setattr__ = dict.__setitem__
    __delattr__ = dict.__delitem__

    def __getstate__(self):
        return self.copy()

    def __setstate__(self, state):
        self.update(state)
