"""On-disk metadata of a snapshot.  The wire format is the reference's (T:manifest.py:30-475):
``.snapshot_metadata`` is ``json.dumps(asdict(SnapshotMetadata), indent=2)`` with keys
``"<rank>/<logical path>"``; field order inside each entry is part of the byte-level contract, which
is why the dataclasses below declare their fields in exactly this order.  Snapshots written by either
implementation are readable by the other."""
from __future__ import annotations

import base64
import json
import struct
from dataclasses import asdict, dataclass, field
from typing import Any, ClassVar, Dict, List, Optional, Tuple, Union

try:  # the YAML fallback reader; same loader preference as the reference (T:manifest.py:22-25)
    from yaml import CSafeLoader as Loader
except ImportError:  # pragma: no cover
    from yaml import SafeLoader as Loader

SNAPSHOT_FORMAT_VERSION = "0.1.0"  # value the reference writes into "version" (T:version.py)


@dataclass
class Entry:
    type: str = field(init=False)
    TYPE: ClassVar[str] = ""

    def __post_init__(self) -> None:
        self.type = self.TYPE

    @classmethod
    def from_yaml_obj(cls, obj: Dict[str, Any]) -> "Entry":
        obj = dict(obj)
        obj.pop("type", None)
        return cls(**obj)


@dataclass
class TensorEntry(Entry):
    TYPE: ClassVar[str] = "Tensor"
    location: str
    serializer: str
    dtype: str
    shape: List[int]
    replicated: bool
    byte_range: Optional[List[int]] = None

    @property
    def byte_range_tuple(self) -> Optional[Tuple[int, int]]:
        return None if self.byte_range is None else (self.byte_range[0], self.byte_range[1])


@dataclass
class Shard:
    offsets: List[int]
    sizes: List[int]
    tensor: TensorEntry

    @classmethod
    def from_yaml_obj(cls, obj: Dict[str, Any]) -> "Shard":
        return cls(offsets=obj["offsets"], sizes=obj["sizes"], tensor=TensorEntry.from_yaml_obj(obj["tensor"]))


def _shards_from(objs: List[Dict[str, Any]]) -> List[Shard]:
    return [Shard.from_yaml_obj(o) for o in objs]


@dataclass
class ShardedTensorEntry(Entry):
    TYPE: ClassVar[str] = "ShardedTensor"
    shards: List[Shard]

    @classmethod
    def from_yaml_obj(cls, obj: Dict[str, Any]) -> "ShardedTensorEntry":
        return cls(shards=_shards_from(obj["shards"]))

    def get_tensor_shape(self) -> List[int]:
        # the furthest corner reached by any shard (T:manifest.py:143-170)
        assert self.shards, "No shards found."
        best = [o + s for o, s in zip(self.shards[0].offsets, self.shards[0].sizes)]
        for sh in self.shards[1:]:
            cand = [o + s for o, s in zip(sh.offsets, sh.sizes)]
            if all(c >= b for c, b in zip(cand, best)):
                best = cand
        return best


@dataclass
class ChunkedTensorEntry(Entry):
    TYPE: ClassVar[str] = "ChunkedTensor"
    dtype: str
    shape: List[int]
    chunks: List[Shard]
    replicated: bool

    @classmethod
    def from_yaml_obj(cls, obj: Dict[str, Any]) -> "ChunkedTensorEntry":
        return cls(dtype=obj["dtype"], shape=obj["shape"], chunks=_shards_from(obj["chunks"]), replicated=obj["replicated"])


NestedList = Union[int, List["NestedList"]]


@dataclass
class DTensorEntry(Entry):
    TYPE: ClassVar[str] = "DTensor"
    shards: List[Shard]
    mesh: NestedList
    dim_map: List[List[int]]

    @classmethod
    def from_yaml_obj(cls, obj: Dict[str, Any]) -> "DTensorEntry":
        return cls(shards=_shards_from(obj["shards"]), mesh=obj["mesh"], dim_map=obj["dim_map"])


@dataclass
class ObjectEntry(Entry):
    TYPE: ClassVar[str] = "object"
    location: str
    serializer: str
    obj_type: str
    replicated: bool


@dataclass
class ListEntry(Entry):
    TYPE: ClassVar[str] = "list"


@dataclass
class DictEntry(Entry):
    TYPE: ClassVar[str] = "dict"
    keys: List[Union[str, int]]


@dataclass
class OrderedDictEntry(Entry):
    TYPE: ClassVar[str] = "OrderedDict"
    keys: List[Union[str, int]]


_PRIMITIVE_TYPES = ("int", "str", "bool", "bytes", "float")


@dataclass
class PrimitiveEntry(Entry):
    """int/str/bool/bytes/float stored inline in the metadata (T:manifest.py:333-420)."""

    supported_types: ClassVar[List[str]] = list(_PRIMITIVE_TYPES)
    serialized_value: str
    replicated: bool
    readable: Optional[str]

    def __init__(self, type: str, serialized_value: str, replicated: bool, readable_value: Optional[str] = None) -> None:
        self.type = type
        self.serialized_value = serialized_value
        self.replicated = replicated
        self.readable = readable_value

    def __post_init__(self) -> None:  # type is an init argument here
        pass

    @staticmethod
    def _encode(kind: str, obj: Any) -> str:
        if kind in ("int", "str", "bool"):
            return str(obj)
        if kind == "bytes":
            return base64.b64encode(obj).decode("utf-8")
        if kind == "float":
            return base64.b64encode(struct.pack("d", float(obj))).decode("utf-8")
        raise TypeError(f"Unsupported primitive obj of type {kind}")

    @classmethod
    def from_object(cls, obj: Any) -> "PrimitiveEntry":
        kind = type(obj).__name__
        if kind not in _PRIMITIVE_TYPES:
            raise TypeError(f"Unsupported primitive obj of type {kind}")
        return cls(kind, cls._encode(kind, obj), False, str(obj) if kind == "float" else None)

    def get_value(self) -> Union[int, str, bool, bytes, float]:
        v = self.serialized_value
        if self.type == "int":
            return int(v)
        if self.type == "str":
            return v
        if self.type == "bool":
            if v not in ("True", "False"):
                raise RuntimeError(f"Unexpected serialized_value for bool type: {v}")
            return v == "True"
        if self.type == "bytes":
            return base64.b64decode(v.encode("utf-8"))
        if self.type == "float":
            return struct.unpack("d", base64.b64decode(v.encode("utf-8")))[0]
        raise ValueError(f"Unable to get deserialized value for {v}")

    @classmethod
    def from_yaml_obj(cls, obj: Dict[str, Any]) -> "PrimitiveEntry":
        if obj["type"] not in _PRIMITIVE_TYPES:
            raise TypeError(f"Unsupported primitive obj of type {obj['type']}")
        return cls(obj["type"], obj["serialized_value"], obj["replicated"], obj.get("readable"))


Manifest = Dict[str, Entry]

_DECODERS = {
    c.TYPE: c
    for c in (TensorEntry, ShardedTensorEntry, ChunkedTensorEntry, DTensorEntry, ObjectEntry, ListEntry, DictEntry, OrderedDictEntry)
}


def entry_from_yaml_obj(obj: Dict[str, Any]) -> Optional[Entry]:
    kind = obj["type"]
    if kind in _PRIMITIVE_TYPES:
        return PrimitiveEntry.from_yaml_obj(obj)
    dec = _DECODERS.get(kind)
    return None if dec is None else dec.from_yaml_obj(obj)


@dataclass
class SnapshotMetadata:
    version: str
    world_size: int
    manifest: Manifest

    def to_yaml(self) -> str:
        # JSON is a subset of YAML; the reference emits JSON for speed (T:manifest.py:442-448)
        return json.dumps(asdict(self), sort_keys=False, indent=2)

    @classmethod
    def from_yaml(cls, yaml_str: str) -> "SnapshotMetadata":
        text = yaml_str
        try:
            d = json.loads(text)
        except ValueError:
            import yaml  # snapshots from very old writers are real YAML

            d = yaml.load(text, Loader=Loader)
        manifest: Manifest = {}
        for path, obj in d["manifest"].items():
            e = entry_from_yaml_obj(obj)
            if e is not None:
                manifest[path] = e
        return cls(version=d["version"], world_size=d["world_size"], manifest=manifest)


# ---- predicates (T:manifest_utils.py) ----------------------------------------------------------
def is_dict_entry(e: Entry) -> bool:
    return isinstance(e, (DictEntry, OrderedDictEntry))


def is_container_entry(e: Entry) -> bool:
    return isinstance(e, (ListEntry, DictEntry, OrderedDictEntry))


def is_fully_replicated_entry(e: Entry) -> bool:
    if isinstance(e, DTensorEntry):
        return all(d[0] == -1 for d in e.dim_map)
    return bool(getattr(e, "replicated", False))


def is_partially_replicated_entry(e: Entry) -> bool:
    if isinstance(e, DTensorEntry):
        n = sum(1 for d in e.dim_map if d[0] == -1)
        return 0 < n < len(e.dim_map)
    return False


def is_replicated_entry(e: Entry) -> bool:
    return is_fully_replicated_entry(e) or is_partially_replicated_entry(e)


def is_sharded_entry(e: Entry) -> bool:
    if isinstance(e, DTensorEntry):
        return any(d[0] != -1 for d in e.dim_map)
    return isinstance(e, ShardedTensorEntry)
