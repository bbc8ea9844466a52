"""Wire format of the gRPC expander (``cluster-autoscaler/expander/grpcplugin/protos/expander.proto``), proto3:

    message BestOptionsRequest  { repeated Option options = 1; map<string, bytes> nodeBytesMap = 3; }
    message BestOptionsResponse { repeated Option options = 1; }
    message Option { string nodeGroupId = 1; int32 nodeCount = 2; string debug = 3; repeated bytes podBytes = 5; }

so that the engine's option vectors (node group id, node count, the pods Estimate() scheduled) can be handed to / taken from an
external expander service byte-compatibly.  Pods and nodes travel as opaque proto-serialized ``v1.Pod`` / ``v1.Node`` bytes
(``grpc_client.go:109-181``): the caller supplies them, this module never looks inside.  No protobuf runtime is needed: the three
messages use only varint and length-delimited fields.  Unknown fields are skipped when decoding (proto3 forward compatibility);
encoding is canonical: fields in number order, map entries sorted by key, default values omitted.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence, Tuple


@dataclass
class Option:
    node_group_id: str = ""
    node_count: int = 0
    debug: str = ""
    pod_bytes: List[bytes] = field(default_factory=list)


@dataclass
class BestOptionsRequest:
    options: List[Option] = field(default_factory=list)
    node_bytes_map: Dict[str, bytes] = field(default_factory=dict)


@dataclass
class BestOptionsResponse:
    options: List[Option] = field(default_factory=list)


# ---- primitives ------------------------------------------------------------------------------------------------------------
def _varint(v: int) -> bytes:
    if v < 0:
        v += 1 << 64          # int32 / int64 negatives are sign-extended to 64 bits (ten bytes)
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _read_varint(buf: bytes, pos: int) -> Tuple[int, int]:
    shift = v = 0
    while True:
        if pos >= len(buf):
            raise ValueError("truncated varint")
        b = buf[pos]
        pos += 1
        v |= (b & 0x7F) << shift
        if not b & 0x80:
            return v, pos
        shift += 7
        if shift > 63:
            raise ValueError("varint too long")


def _len_field(num: int, payload: bytes) -> bytes:
    return _varint(num << 3 | 2) + _varint(len(payload)) + payload


def _fields(buf: bytes):
    """Yield (field number, wire type, value) with value = int for varint / fixed, bytes for length-delimited."""
    pos = 0
    while pos < len(buf):
        key, pos = _read_varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == 0:
            v, pos = _read_varint(buf, pos)
        elif wt == 2:
            n, pos = _read_varint(buf, pos)
            if pos + n > len(buf):
                raise ValueError("truncated length-delimited field")
            v, pos = buf[pos:pos + n], pos + n
        elif wt == 1:
            v, pos = int.from_bytes(buf[pos:pos + 8], "little"), pos + 8
        elif wt == 5:
            v, pos = int.from_bytes(buf[pos:pos + 4], "little"), pos + 4
        else:
            raise ValueError("unsupported wire type %d" % wt)
        yield num, wt, v


# ---- messages --------------------------------------------------------------------------------------------------------------
def encode_option(o: Option) -> bytes:
    out = b""
    if o.node_group_id:
        out += _len_field(1, o.node_group_id.encode("utf-8"))
    if o.node_count:
        out += _varint(2 << 3 | 0) + _varint(int(o.node_count))
    if o.debug:
        out += _len_field(3, o.debug.encode("utf-8"))
    for pb in o.pod_bytes:
        out += _len_field(5, bytes(pb))
    return out


def decode_option(buf: bytes) -> Option:
    o = Option()
    for num, wt, v in _fields(buf):
        if num == 1 and wt == 2:
            o.node_group_id = v.decode("utf-8")
        elif num == 2 and wt == 0:
            v &= 0xFFFFFFFF
            o.node_count = v - (1 << 32) if v & 0x80000000 else v
        elif num == 3 and wt == 2:
            o.debug = v.decode("utf-8")
        elif num == 5 and wt == 2:
            o.pod_bytes.append(bytes(v))
    return o


def encode_request(r: BestOptionsRequest) -> bytes:
    out = b"".join(_len_field(1, encode_option(o)) for o in r.options)
    for k in sorted(r.node_bytes_map):
        entry = _len_field(1, k.encode("utf-8")) + _len_field(2, bytes(r.node_bytes_map[k]))   # map entries always carry key and value
        out += _len_field(3, entry)
    return out


def decode_request(buf: bytes) -> BestOptionsRequest:
    r = BestOptionsRequest()
    for num, wt, v in _fields(buf):
        if num == 1 and wt == 2:
            r.options.append(decode_option(v))
        elif num == 3 and wt == 2:
            k, val = "", b""
            for n2, w2, v2 in _fields(v):
                if n2 == 1 and w2 == 2:
                    k = v2.decode("utf-8")
                elif n2 == 2 and w2 == 2:
                    val = bytes(v2)
            r.node_bytes_map[k] = val
    return r


def encode_response(r: BestOptionsResponse) -> bytes:
    return b"".join(_len_field(1, encode_option(o)) for o in r.options)


def decode_response(buf: bytes) -> BestOptionsResponse:
    return BestOptionsResponse([decode_option(v) for num, wt, v in _fields(buf) if num == 1 and wt == 2])


# ---- the client's bookkeeping (grpc_client.go:109-158) -----------------------------------------------------------------------
def populate_options_for_grpc(options: Sequence[Tuple[str, int, str, Sequence[bytes]]]) -> Tuple[List[Option], Dict[str, int]]:
    """(node group id, node count, debug, pod bytes) per expansion option -> gRPC options + node group id -> option index."""
    msgs, by_id = [], {}
    for i, (ng, count, debug, pods) in enumerate(options):
        by_id[ng] = i
        msgs.append(Option(ng, int(count), debug, [bytes(p) for p in pods]))
    return msgs, by_id


def transform_and_sanitize_options_from_grpc(response: Optional[BestOptionsResponse], by_id: Dict[str, int]) -> Optional[List[int]]:
    """Indices of the caller's options the server picked, unknown node group ids dropped (grpc_client.go:140-156);
    None = nil / empty response (:95-98)."""
    if response is None or not response.options:
        return None
    return [by_id[o.node_group_id] for o in response.options if o.node_group_id in by_id]
