"""gRPC expander wire format (expander/grpcplugin/protos/expander.proto): the hand-written codec must be byte-compatible with the
protobuf runtime on the same schema (built from a descriptor here, nothing generated is checked in), round-trip like
protos/round_trip_test.go:31-118 (opaque pod / node bytes survive), and sanitize responses like grpc_client.go:140-156."""
import random

import pytest

from kubernetes_autoscaler_b200 import grpcwire as gw


def _runtime_messages():
    descriptor_pb2 = pytest.importorskip("google.protobuf.descriptor_pb2")
    from google.protobuf import descriptor_pool, message_factory
    f = descriptor_pb2.FileDescriptorProto(name="expander_test.proto", package="grpcplugin", syntax="proto3")
    T = descriptor_pb2.FieldDescriptorProto
    opt = f.message_type.add(name="Option")
    opt.field.add(name="nodeGroupId", number=1, type=T.TYPE_STRING, label=T.LABEL_OPTIONAL)
    opt.field.add(name="nodeCount", number=2, type=T.TYPE_INT32, label=T.LABEL_OPTIONAL)
    opt.field.add(name="debug", number=3, type=T.TYPE_STRING, label=T.LABEL_OPTIONAL)
    opt.field.add(name="podBytes", number=5, type=T.TYPE_BYTES, label=T.LABEL_REPEATED)
    req = f.message_type.add(name="BestOptionsRequest")
    req.field.add(name="options", number=1, type=T.TYPE_MESSAGE, label=T.LABEL_REPEATED, type_name=".grpcplugin.Option")
    entry = req.nested_type.add(name="NodeBytesMapEntry")
    entry.options.map_entry = True
    entry.field.add(name="key", number=1, type=T.TYPE_STRING, label=T.LABEL_OPTIONAL)
    entry.field.add(name="value", number=2, type=T.TYPE_BYTES, label=T.LABEL_OPTIONAL)
    req.field.add(name="nodeBytesMap", number=3, type=T.TYPE_MESSAGE, label=T.LABEL_REPEATED,
                  type_name=".grpcplugin.BestOptionsRequest.NodeBytesMapEntry")
    resp = f.message_type.add(name="BestOptionsResponse")
    resp.field.add(name="options", number=1, type=T.TYPE_MESSAGE, label=T.LABEL_REPEATED, type_name=".grpcplugin.Option")
    pool = descriptor_pool.DescriptorPool()
    pool.Add(f)
    get = lambda n: message_factory.GetMessageClass(pool.FindMessageTypeByName("grpcplugin." + n))
    return get("Option"), get("BestOptionsRequest"), get("BestOptionsResponse")


def _rand_request(rng):
    opts = []
    for i in range(rng.randint(0, 5)):
        opts.append(gw.Option("ng-%d" % rng.randint(0, 9) if rng.random() < 0.9 else "", rng.choice([0, 1, 3, 127, 128, 300, 2 ** 31 - 1, -1]),
                              rng.choice(["", "dbg", "least-waste: 0.5 ✓"]),
                              [bytes(rng.getrandbits(8) for _ in range(rng.randint(0, 40))) for _ in range(rng.randint(0, 3))]))
    nodes = {"node-%d" % i: bytes(rng.getrandbits(8) for _ in range(rng.randint(0, 60))) for i in range(rng.randint(0, 4))}
    return gw.BestOptionsRequest(opts, nodes)


def test_codec_matches_the_protobuf_runtime():
    Option, Request, Response = _runtime_messages()
    rng = random.Random(5)
    for _ in range(200):
        r = _rand_request(rng)
        m = Request()
        for o in r.options:
            mo = m.options.add(nodeGroupId=o.node_group_id, nodeCount=o.node_count, debug=o.debug)
            mo.podBytes.extend(o.pod_bytes)
        for k, v in r.node_bytes_map.items():
            m.nodeBytesMap[k] = v
        assert gw.encode_request(r) == m.SerializeToString(deterministic=True)          # byte-identical canonical encoding
        back = gw.decode_request(m.SerializeToString())
        assert back == r
        m2 = Request()
        m2.ParseFromString(gw.encode_request(r))
        assert m2 == m
        resp = gw.BestOptionsResponse(r.options[:2])
        mr = Response()
        mr.ParseFromString(gw.encode_response(resp))
        assert [(o.nodeGroupId, o.nodeCount, o.debug, list(o.podBytes)) for o in mr.options] == \
               [(o.node_group_id, o.node_count, o.debug, o.pod_bytes) for o in resp.options]
        assert gw.decode_response(mr.SerializeToString()) == resp


def test_round_trip_like_the_reference():
    """protos/round_trip_test.go: a request with one option holding pod bytes and a node-bytes map survives; so does a response."""
    pod_bytes, node_bytes = b"\x0a\x0c\x0a\x04test\x12\x04test" + bytes(range(40)), b"\x0a\x0c\x0a\x04test\x12\x04test\x1a\x00"
    r = gw.BestOptionsRequest([gw.Option(pod_bytes=[pod_bytes])], {"node": node_bytes})
    r2 = gw.decode_request(gw.encode_request(r))
    assert r2 == r and r2.options[0].pod_bytes[0] == pod_bytes and r2.node_bytes_map["node"] == node_bytes
    resp = gw.BestOptionsResponse([gw.Option("ng1", 2, "d", [pod_bytes])])
    assert gw.decode_response(gw.encode_response(resp)) == resp
    assert gw.decode_request(b"") == gw.BestOptionsRequest() and gw.encode_request(gw.BestOptionsRequest()) == b""
    # unknown fields (a newer server) are skipped
    assert gw.decode_option(gw.encode_option(gw.Option("a", 1)) + b"\x48\x07" + b"\x52\x02hi") == gw.Option("a", 1)
    with pytest.raises(ValueError):
        gw.decode_option(b"\x0a\x05ab")


def test_sanitize_like_the_client():
    """grpc_client.go:140-156: options come back by node group id, unknown ids are dropped, an empty response means nil."""
    msgs, by_id = gw.populate_options_for_grpc([("ng1", 2, "", []), ("ng2", 1, "", [b"p"]), ("ng3", 4, "x", [])])
    assert [m.node_group_id for m in msgs] == ["ng1", "ng2", "ng3"] and msgs[1].pod_bytes == [b"p"]
    wire = gw.encode_response(gw.BestOptionsResponse([gw.Option("ng3", 4), gw.Option("bogus", 1), gw.Option("ng1", 2)]))
    assert gw.transform_and_sanitize_options_from_grpc(gw.decode_response(wire), by_id) == [2, 0]
    assert gw.transform_and_sanitize_options_from_grpc(gw.BestOptionsResponse(), by_id) is None
    assert gw.transform_and_sanitize_options_from_grpc(None, by_id) is None
