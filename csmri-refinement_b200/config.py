"""JSON run configurations - mirror of utils/config.py, as far as the hot path's
callers use it.

The reference's ``configs/1-recnet.json`` and ``configs/2-refinement.json`` are
read *unchanged*; this module gives them the attribute-bag object the
reference's builders expect (``utils/config.py:36-250``):

* top-level keys become attributes, nested objects stay plain dicts until a
  builder wraps one with ``Configuration.from_dict(section, parent)``
  (training/runner.py:19, training/adversarial_runner.py:26-36);
* ``"seed"`` is stored as ``_seed`` / ``.seed`` (utils/config.py:22-24,50-52);
* ``"#include"`` (merge a file into the object, :10-19) and a top-level
  ``"include": {key: path}`` section (:236-250) are resolved relative to the
  including file;
* ``get_attr(key, default, alternative)``, ``has_attr``, ``to_param_dict(required,
  optional, key_renames)`` (:71-93,151-184) and the ``--conf key=value`` string
  conversion of ``update`` (:108-149) behave as in the reference.
"""
import json
import os

TYPE_TAG = '__type__'


class Configuration(object):
    def __init__(self):
        self._seed = 0
        self._src_file = None
        self.__dict__[TYPE_TAG] = str(type(self))

    # -- construction ------------------------------------------------------
    @staticmethod
    def from_dict(dictionary, parent_config=None):
        """Wrap a dict (e.g. the ``model`` section); seed and source file are
        inherited from ``parent_config``."""
        if isinstance(dictionary, Configuration):
            return dictionary
        conf = Configuration()
        conf.__dict__.update(dictionary)
        if parent_config is not None:
            conf._seed = parent_config._seed
            conf._src_file = parent_config._src_file
        return conf

    @staticmethod
    def from_json(src):
        base = os.path.dirname(src)

        def resolve(path):
            return path if os.path.isabs(path) else os.path.join(base, path)

        def hook(obj):
            merged = {}
            inc = obj.pop('#include', None)
            if inc is not None:
                for path in (inc if isinstance(inc, list) else [inc]):
                    merged.update(Configuration.from_json(resolve(path)).__dict__)
            if 'seed' in obj:
                merged['_seed'] = obj.pop('seed')
            merged.update(obj)
            if obj.get(TYPE_TAG) == str(Configuration):
                return Configuration.from_dict(merged)
            return merged

        with open(src, 'r') as f:
            conf = json.load(f, object_hook=hook)
        conf = Configuration.from_dict(conf)
        conf._src_file = src
        includes = conf.__dict__.pop('include', None)
        if includes:
            for key, path in includes.items():
                sub = Configuration.from_json(resolve(path)).__dict__
                if key == '':
                    conf.__dict__ = dict(sub, **conf.__dict__)
                else:
                    own = conf.__dict__.get(key)
                    conf.__dict__[key] = dict(sub)
                    if isinstance(own, dict):
                        conf.__dict__[key].update(own)
        return conf

    # -- access --------------------------------------------------------------
    @property
    def seed(self):
        return self._seed

    @property
    def file(self):
        return self._src_file

    def has_attr(self, key):
        return hasattr(self, key)

    def get_attr(self, key, default=None, alternative=None):
        if hasattr(self, key):
            return getattr(self, key)
        if alternative is None:
            return default
        value = self.get_attr(alternative)
        if value is None:
            raise ValueError('Configuration did not contain {} or alternative {}'.format(
                key, alternative))
        return value

    def to_param_dict(self, required_params=(), optional_params=(), key_renames=None):
        key_renames = key_renames or {}
        params = {}
        for key in required_params:
            value = self.get_attr(key)
            assert value is not None, 'Parameter {} is marked as required'.format(key)
            params[key] = value
        if isinstance(optional_params, dict):
            for key, default in optional_params.items():
                params[key] = self.get_attr(key, default=default)
        else:
            for key in optional_params:
                value = self.get_attr(key)
                if value is not None:
                    params[key] = value
        return {key_renames.get(k, k): v for k, v in params.items()}

    def update(self, values_by_keys):
        """``--conf key=value`` overrides: strings are converted to bool / int /
        float / flat list where they parse as such."""
        def convert(s):
            if (s.startswith('[') and s.endswith(']')) or (s.startswith('(') and s.endswith(')')):
                return [convert(e.strip()) for e in s[1:-1].split(',')]
            if s in ('True', 'False'):
                return s == 'True'
            for cast in (int, float):
                try:
                    return cast(s)
                except ValueError:
                    pass
            return s

        for key, value in values_by_keys.items():
            self.__dict__['_seed' if key == 'seed' else key] = convert(value)

    def serialize(self, dst):
        with open(dst, 'w') as f:
            json.dump(self.__dict__, f, default=lambda o: o.__dict__, indent=2)

    def __str__(self):
        return 'Configuration object\n' + ''.join(
            '  {}: {}\n'.format(k, v) for k, v in self.__dict__.items())
