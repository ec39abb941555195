"""TEST INFRASTRUCTURE ONLY -- `Data` is imported but never used on the hot path."""


class Data:  # pragma: no cover
    def __init__(self, **kw):
        self.__dict__.update(kw)
