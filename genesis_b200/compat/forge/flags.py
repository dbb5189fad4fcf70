"""Minimal Forge-compatible flag registry (reference: forge/forge/flags.py:27-132).

Same public surface -- DEFINE_{string,integer,boolean,bool,float} and FLAGS -- on a global
argparse parser: plugin files register their flags at import time, `FLAGS.<name>` parses the
command line lazily, booleans accept --flag, --flag=True and --noflag.  Written for this repo;
no TensorFlow dependency."""
import argparse as _argparse

_parser = _argparse.ArgumentParser(allow_abbrev=False)
_defined = {}


class _FlagValues(object):
    def __init__(self):
        self.__dict__['__flags'] = {}
        self.__dict__['__parsed'] = False

    def _parse_flags(self, args=None):
        known, unparsed = _parser.parse_known_args(args=args)
        self.__dict__['__flags'].update(vars(known))
        self.__dict__['__parsed'] = True
        return unparsed

    def _ensure(self):
        if not self.__dict__['__parsed']:
            self._parse_flags()

    def __getattr__(self, name):
        self._ensure()
        try:
            return self.__dict__['__flags'][name]
        except KeyError:
            raise AttributeError(name)

    def __setattr__(self, name, value):
        self._ensure()
        self.__dict__['__flags'][name] = value

    # models also use item access on cfg (monet_config.py:56-57)
    def __getitem__(self, name):
        return getattr(self, name)

    def __setitem__(self, name, value):
        setattr(self, name, value)

    def __contains__(self, name):
        self._ensure()
        return name in self.__dict__['__flags']


FLAGS = _FlagValues()


def _define(name, default, doc, kind):
    if name in _defined:          # re-import of a config file: keep the first definition
        return
    _defined[name] = default
    _parser.add_argument('--' + name, default=default, help=doc, type=kind)


def DEFINE_string(name, default, doc):
    _define(name, default, doc, str)


def DEFINE_integer(name, default, doc):
    _define(name, default, doc, int)


def DEFINE_float(name, default, doc):
    _define(name, default, doc, float)


def DEFINE_boolean(name, default, doc):
    if name in _defined:
        return
    _defined[name] = default
    _parser.add_argument('--' + name, nargs='?', const=True, default=default, help=doc,
                         type=lambda v: v.lower() in ('true', 't', '1'))
    _parser.add_argument('--no' + name, action='store_false', dest=name.replace('-', '_'))


DEFINE_bool = DEFINE_boolean


def defaults():
    """name -> default of every flag registered so far (used to build a cfg without argv)."""
    return dict(_defined)
