"""Stand-in for the `attrdict` package the reference imports (train.py:19,
models/*_config.py): a dict whose keys are also attributes.  `AttrDefault(factory, {})`
is used at train.py:498 as a defaultdict with attribute access."""


class AttrDict(dict):
    def __getattr__(self, name):
        try:
            return self[name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self[name] = value

    def __delattr__(self, name):
        try:
            del self[name]
        except KeyError:
            raise AttributeError(name)


class AttrDefault(AttrDict):
    def __init__(self, default_factory=None, items=None):
        super().__init__(items or {})
        dict.__setattr__(self, '_factory', default_factory)

    def __missing__(self, key):
        if self._factory is None:
            raise KeyError(key)
        self[key] = value = self._factory()
        return value

    def __getattr__(self, name):
        if name.startswith('_'):
            raise AttributeError(name)
        return self[name]
