"""Reversible flattening of nested state dicts into ``logical_path -> leaf`` maps.
Path grammar is the reference's (T:flatten.py:20-226): components joined by "/", user keys
percent-encoded ("%" -> %25, "/" -> %2F); lists and str/int-keyed dicts are containers, anything else
is a leaf."""
from __future__ import annotations

from collections import OrderedDict
from typing import Any, Dict, List, Tuple
from urllib.parse import unquote

from .manifest import DictEntry, Entry, ListEntry, Manifest, OrderedDictEntry


def _encode(s: str) -> str:
    return s.replace("%", "%25").replace("/", "%2F")


def _decode(s: str) -> str:
    return unquote(s)


def _flattenable_dict(d: Dict[Any, Any]) -> bool:
    keys = list(d.keys())
    if any(not isinstance(k, (str, int)) for k in keys):
        return False
    return len({str(k) for k in keys}) == len(keys)


def flatten(obj: Any, prefix: str) -> Tuple[Manifest, Dict[str, Any]]:
    manifest: Manifest = {}
    leaves: Dict[str, Any] = {}
    # explicit stack, children pushed in reverse so that pop order == insertion order
    stack: List[Tuple[str, Any]] = [(_encode(prefix), obj)]
    while stack:
        path, node = stack.pop()
        kind = type(node)
        if kind is list:
            manifest[path] = ListEntry()
            stack.extend((f"{path}/{i}", node[i]) for i in range(len(node) - 1, -1, -1))
        elif kind in (dict, OrderedDict) and _flattenable_dict(node):
            keys = list(node.keys())
            manifest[path] = DictEntry(keys=keys) if kind is dict else OrderedDictEntry(keys=keys)
            stack.extend((f"{path}/{_encode(str(k))}", node[k]) for k in reversed(keys))
        else:
            leaves[path] = node
    return manifest, leaves


def _is_int(s: str) -> bool:
    return s.isdigit() or (len(s) > 1 and s[0] in "+-" and s[1:].isdigit())


def inflate(manifest: Manifest, flattened: Dict[str, Any], prefix: str) -> Any:
    root = _encode(prefix)
    manifest = {k: v for k, v in manifest.items() if k.split("/", 1)[0] == root}
    flattened = {k: v for k, v in flattened.items() if k.split("/", 1)[0] == root}
    if root in flattened:
        return flattened[root]
    if root not in manifest:
        raise AssertionError(f"{root} is absent in both manifest and flattened.")

    def make(e: Entry) -> Any:
        if isinstance(e, ListEntry):
            return []
        if isinstance(e, DictEntry):
            return dict.fromkeys(e.keys)
        if isinstance(e, OrderedDictEntry):
            return OrderedDict.fromkeys(e.keys)
        raise RuntimeError(f"Unrecognized container entry type: {type(e)} ({e.type}).")

    containers = {p: make(e) for p, e in manifest.items()}
    children: Dict[str, Dict[str, Any]] = {}
    for path, val in list(containers.items()) + list(flattened.items()):
        if path == root:
            continue
        parent, _, key = path.rpartition("/")
        children.setdefault(parent, {})[key] = val
    for path, kids in children.items():
        c = containers[path]
        if isinstance(c, list):
            c.extend(v for _, v in sorted(kids.items(), key=lambda kv: int(kv[0])))
        else:
            by_key: Dict[Any, Any] = {}
            for k, v in kids.items():
                dk = _decode(k)
                by_key[dk] = v
                if _is_int(dk):
                    by_key[int(dk)] = v
            # keys recorded in the entry but absent from the data are dropped (T:flatten.py:189-196)
            for k in list(c.keys()):
                if k in by_key:
                    c[k] = by_key[k]
                else:
                    del c[k]
    return containers[root]
