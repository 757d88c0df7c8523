"""Minimal yacs-style ``CfgNode``: attribute access, yaml files with ``_BASE_``
inheritance, ``merge_from_list(["KEY.SUB", value, ...])``, ``freeze`` / ``clone``.
Mirrors the parts of d2!/config/config.py + fvcore CfgNode that
``fsod_train_net.py:76-89`` (setup) exercises."""
from __future__ import annotations

import ast
import copy
import os
from typing import Any, Dict, List

import yaml

_BASE_KEY = "_BASE_"


class CfgNode(dict):
    def __init__(self, init_dict: Dict[str, Any] = None, key_list=None, new_allowed: bool = False):
        super().__init__()
        self.__dict__["_frozen"] = False
        self.__dict__["_new_allowed"] = new_allowed
        for k, v in (init_dict or {}).items():
            self[k] = CfgNode(v) if isinstance(v, dict) and not isinstance(v, CfgNode) else v

    # ---- attribute access
    def __getattr__(self, name: str) -> Any:
        if name in self:
            return self[name]
        raise AttributeError(name)

    def __setattr__(self, name: str, value: Any) -> None:
        if self.__dict__.get("_frozen", False):
            raise AttributeError(f"Attempted to set {name} to {value}, but CfgNode is immutable")
        self[name] = value

    # ---- state
    def freeze(self) -> None:
        self._set_frozen(True)

    def defrost(self) -> None:
        self._set_frozen(False)

    def is_frozen(self) -> bool:
        return self.__dict__["_frozen"]

    def _set_frozen(self, flag: bool) -> None:
        self.__dict__["_frozen"] = flag
        for v in self.values():
            if isinstance(v, CfgNode):
                v._set_frozen(flag)

    def clone(self) -> "CfgNode":
        return copy.deepcopy(self)

    def __deepcopy__(self, memo):
        out = CfgNode()
        for k, v in self.items():
            dict.__setitem__(out, k, copy.deepcopy(v, memo))
        out.__dict__["_frozen"] = self.__dict__["_frozen"]
        return out

    # ---- merging
    @classmethod
    def load_yaml_with_base(cls, filename: str, allow_unsafe: bool = False) -> Dict[str, Any]:
        with open(filename, "r") as f:
            cfg = yaml.safe_load(f) or {}

        def merge_a_into_b(a: Dict[str, Any], b: Dict[str, Any]) -> None:
            for k, v in a.items():
                if isinstance(v, dict) and isinstance(b.get(k), dict):
                    merge_a_into_b(v, b[k])
                else:
                    b[k] = v

        if _BASE_KEY in cfg:
            base = cfg.pop(_BASE_KEY)
            if base.startswith("~"):
                base = os.path.expanduser(base)
            if not os.path.isabs(base):
                base = os.path.join(os.path.dirname(filename), base)
            base_cfg = cls.load_yaml_with_base(base, allow_unsafe)
            merge_a_into_b(cfg, base_cfg)
            return base_cfg
        return cfg

    def merge_from_file(self, cfg_filename: str, allow_unsafe: bool = True) -> None:
        if not os.path.isfile(cfg_filename):
            raise FileNotFoundError(f"Config file '{cfg_filename}' does not exist!")
        self.merge_from_other_cfg(CfgNode(self.load_yaml_with_base(cfg_filename, allow_unsafe)))

    def merge_from_other_cfg(self, other: "CfgNode") -> None:
        _merge(other, self, [])

    def merge_from_list(self, cfg_list: List[Any]) -> None:
        if len(cfg_list) % 2:
            raise ValueError(f"Override list has odd length: {cfg_list}; it must be a list of pairs")
        for full_key, v in zip(cfg_list[0::2], cfg_list[1::2]):
            node = self
            parts = full_key.split(".")
            for p in parts[:-1]:
                if p not in node:
                    raise KeyError(f"Non-existent key: {full_key}")
                node = node[p]
            if parts[-1] not in node:
                raise KeyError(f"Non-existent key: {full_key}")
            node[parts[-1]] = _coerce(_decode(v), node[parts[-1]], full_key)

    def dump(self, **kwargs) -> str:
        def plain(n):
            return {k: plain(v) for k, v in n.items()} if isinstance(n, CfgNode) else (list(n) if isinstance(n, tuple) else n)
        return yaml.safe_dump(plain(self), **kwargs)


def _decode(v: Any) -> Any:
    if not isinstance(v, str):
        return v
    try:
        return ast.literal_eval(v)
    except (ValueError, SyntaxError):
        return v


def _coerce(new: Any, old: Any, key: str) -> Any:
    if old is None or new is None or type(new) is type(old):
        return new
    if isinstance(old, float) and isinstance(new, int) and not isinstance(new, bool):
        return float(new)
    if isinstance(old, (list, tuple)) and isinstance(new, (list, tuple)):
        return type(old)(new)
    if isinstance(old, str) and not isinstance(new, str):
        return new if not isinstance(new, (int, float)) else str(new)
    raise ValueError(f"Type mismatch ({type(old)} vs. {type(new)}) for config key: {key}")


def _merge(a: CfgNode, b: CfgNode, stack: List[str]) -> None:
    for k, v in a.items():
        full = ".".join(stack + [k])
        if k not in b:
            if b.__dict__.get("_new_allowed", False):
                dict.__setitem__(b, k, copy.deepcopy(v))
                continue
            raise KeyError(f"Non-existent config key: {full}")
        if isinstance(v, CfgNode) and isinstance(b[k], CfgNode):
            _merge(v, b[k], stack + [k])
        else:
            # yaml leaves "(1, 2)" as a string; decode it like yacs does
            dict.__setitem__(b, k, _coerce(_decode(copy.deepcopy(v)), b[k], full))
