"""BODY_HEAD_REGISTRY, mirror of regressor/human_shape/models/body_heads/registry.py:1-7 (the reference
uses fvcore's Registry; a dict-backed equivalent avoids the dependency)."""


class Registry(dict):
    def __init__(self, name):
        super().__init__()
        self._name = name

    def register(self, obj=None):
        if obj is None:
            def deco(cls):
                self[cls.__name__] = cls
                return cls
            return deco
        self[obj.__name__] = obj
        return obj


BODY_HEAD_REGISTRY = Registry('BODY_HEAD_REGISTRY')
