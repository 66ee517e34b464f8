"""Minimal stand-in for yacs.config.CfgNode (yacs is the reference's config dependency, IST/config/defaults.py:2; it is
not installed in the build image). Attribute access, clone, freeze/defrost, merge_from_file (YAML) and merge_from_list."""
import copy


class CfgNode(dict):
    _IMMUTABLE = "__immutable__"

    def __init__(self, init=None):
        super().__init__()
        self.__dict__[CfgNode._IMMUTABLE] = False
        for k, v in (init or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    def __getattr__(self, name):
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name, value):
        if self.__dict__[CfgNode._IMMUTABLE]:
            raise AttributeError(f"Attempted to set {name} to {value}, but CfgNode is immutable")
        self[name] = value

    def is_frozen(self):
        return self.__dict__[CfgNode._IMMUTABLE]

    def _set_immutable(self, flag):
        self.__dict__[CfgNode._IMMUTABLE] = flag
        for v in self.values():
            if isinstance(v, CfgNode):
                v._set_immutable(flag)

    def freeze(self):
        self._set_immutable(True)

    def defrost(self):
        self._set_immutable(False)

    def clone(self):
        out = CfgNode()
        for k, v in self.items():
            out[k] = v.clone() if isinstance(v, CfgNode) else copy.deepcopy(v)
        return out

    def _merge(self, other, path=()):
        for k, v in other.items():
            if k not in self:
                raise KeyError("Non-existent config key: " + ".".join(path + (k,)))
            if isinstance(self[k], CfgNode) and isinstance(v, dict):
                self[k]._merge(v, path + (k,))
            else:
                self[k] = v

    def merge_from_file(self, cfg_filename):
        import yaml
        with open(cfg_filename, "r") as f:
            self._merge(yaml.safe_load(f) or {})

    def merge_from_list(self, cfg_list):
        import ast
        assert len(cfg_list) % 2 == 0, "Override list has odd length"
        for full_key, v in zip(cfg_list[0::2], cfg_list[1::2]):
            node = self
            keys = full_key.split(".")
            for k in keys[:-1]:
                node = node[k]
            if keys[-1] not in node:
                raise KeyError("Non-existent config key: " + full_key)
            if isinstance(v, str):
                try:
                    v = ast.literal_eval(v)
                except (ValueError, SyntaxError):
                    pass
            node[keys[-1]] = v
