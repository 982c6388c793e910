"""The hand-rolled protobuf codec (common_b200/wire.py) against the real protobuf runtime, with descriptors built
from microscopes/io/schema.proto:3-55 (no protoc here: the messages are declared programmatically)."""
import numpy as np
import pytest

from common_b200 import wire

pb = pytest.importorskip("google.protobuf")
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory  # noqa: E402

F = descriptor_pb2.FieldDescriptorProto


def _pool():
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name, fd.package, fd.syntax = "schema.proto", "microscopes.io", "proto2"

    def msg(name, fields):
        m = fd.message_type.add()
        m.name = name
        for fname, num, ftype, label, tname in fields:
            f = m.field.add()
            f.name, f.number, f.type, f.label = fname, num, ftype, label
            if tname:
                f.type_name = ".microscopes.io." + tname
    REQ, REP = F.LABEL_REQUIRED, F.LABEL_REPEATED
    msg("CRP", [("alpha", 1, F.TYPE_FLOAT, REQ, None)])
    msg("GroupData", [("id", 1, F.TYPE_UINT32, REQ, None), ("data", 2, F.TYPE_BYTES, REQ, None)])
    msg("GroupManager", [("alpha", 1, F.TYPE_FLOAT, REQ, None), ("assignments", 2, F.TYPE_INT32, REP, None),
                         ("groups", 3, F.TYPE_MESSAGE, REP, "GroupData")])
    msg("MixtureModelGroup", [("suffstats", 1, F.TYPE_BYTES, REP, None)])
    msg("MixtureModelState", [("hypers", 1, F.TYPE_BYTES, REP, None), ("groups", 2, F.TYPE_BYTES, REQ, None)])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return pool


def _cls(pool, name):
    return message_factory.GetMessageClass(pool.FindMessageTypeByName("microscopes.io." + name))


def test_group_manager_bytes_equal_the_protobuf_runtime():
    pool = _pool()
    rng = np.random.default_rng(3)
    assign = rng.integers(-1, 300, size=5000).tolist()
    groups = [{"id": int(g), "data": bytes(rng.integers(0, 256, size=int(rng.integers(0, 40)), dtype=np.uint8))} for g in range(0, 300, 7)]
    mine = wire.encode("GroupManager", {"alpha": 0.75, "assignments": assign, "groups": groups})
    ref = _cls(pool, "GroupManager")()
    ref.alpha = 0.75
    ref.assignments.extend(assign)
    for g in groups:
        x = ref.groups.add()
        x.id, x.data = g["id"], g["data"]
    assert mine == ref.SerializeToString()
    back = wire.decode("GroupManager", ref.SerializeToString())
    assert back["assignments"] == assign and back["alpha"] == 0.75
    assert [(g["id"], g["data"]) for g in back["groups"]] == [(g["id"], g["data"]) for g in groups]


def test_state_and_group_messages_round_trip_through_the_runtime():
    pool = _pool()
    inner = wire.encode("MixtureModelGroup", {"suffstats": [b"\x01\x02", b"", b"xyz"]})
    g = _cls(pool, "MixtureModelGroup")()
    g.ParseFromString(inner)
    assert list(g.suffstats) == [b"\x01\x02", b"", b"xyz"] and g.SerializeToString() == inner
    top = wire.encode("MixtureModelState", {"hypers": [b"h0", b"h1"], "groups": inner})
    s = _cls(pool, "MixtureModelState")()
    s.ParseFromString(top)
    assert list(s.hypers) == [b"h0", b"h1"] and s.groups == inner and s.SerializeToString() == top
    assert wire.decode("CRP", wire.encode("CRP", {"alpha": 2.5})) == {"alpha": 2.5}


def test_packed_repeated_scalars_and_unknown_fields_are_accepted():
    # proto3-style writers pack repeated scalars; unknown fields are skipped
    import struct
    packed_counts = b"\x0a\x03\x05\x00\x07"                       # field 1, length-delimited: 5, 0, 7
    assert wire.decode("dd.Group", packed_counts + b"\x78\x01")["counts"] == [5, 0, 7]   # + unknown field 15
    packed_alphas = b"\x0a\x08" + struct.pack("<ff", 0.5, 1.5)
    assert wire.decode("dd.Shared", packed_alphas)["alphas"] == [0.5, 1.5]
    msg = wire.encode("nich.Group", {"count": 7, "mean": 0.25, "count_times_variance": 3.5})
    assert wire.decode("nich.Group", msg) == {"count": 7, "mean": 0.25, "count_times_variance": 3.5}


def test_in_tree_model_messages_equal_the_protobuf_runtime():
    # BetaBernoulliNonConj.{Shared, Group} and DirichletMultinomial.{Shared, Group} (schema.proto:7-29): the bytes of
    # group::get_ss / hypers::get_hp for the two in-tree models
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name, fd.package, fd.syntax = "schema_models.proto", "microscopes.io", "proto2"
    REQ, REP = F.LABEL_REQUIRED, F.LABEL_REPEATED

    def nested(outer, name, fields):
        m = outer.nested_type.add()
        m.name = name
        for fname, num, ftype, label in fields:
            f = m.field.add()
            f.name, f.number, f.type, f.label = fname, num, ftype, label

    bbnc = fd.message_type.add(); bbnc.name = "BetaBernoulliNonConj"
    nested(bbnc, "Shared", [("alpha", 1, F.TYPE_FLOAT, REQ), ("beta", 2, F.TYPE_FLOAT, REQ)])
    nested(bbnc, "Group", [("p", 1, F.TYPE_FLOAT, REQ), ("heads", 2, F.TYPE_UINT32, REQ), ("tails", 3, F.TYPE_UINT32, REQ)])
    dm = fd.message_type.add(); dm.name = "DirichletMultinomial"
    nested(dm, "Shared", [("alphas", 1, F.TYPE_FLOAT, REP)])
    nested(dm, "Group", [("counts", 1, F.TYPE_UINT32, REP), ("ratio", 2, F.TYPE_FLOAT, REQ)])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)

    def cls(name):
        return message_factory.GetMessageClass(pool.FindMessageTypeByName("microscopes.io." + name))

    s = cls("BetaBernoulliNonConj.Shared")(); s.alpha, s.beta = 0.5, 2.0
    assert wire.encode("bbnc.Shared", {"alpha": 0.5, "beta": 2.0}) == s.SerializeToString()
    g = cls("BetaBernoulliNonConj.Group")(); g.p, g.heads, g.tails = 0.3125, 7, 300
    mine = wire.encode("bbnc.Group", {"p": 0.3125, "heads": 7, "tails": 300})
    assert mine == g.SerializeToString() and wire.decode("bbnc.Group", mine) == {"p": 0.3125, "heads": 7, "tails": 300}
    s = cls("DirichletMultinomial.Shared")(); s.alphas.extend([1.0, 0.25, 3.5])
    assert wire.encode("dm.Shared", {"alphas": [1.0, 0.25, 3.5]}) == s.SerializeToString()
    g = cls("DirichletMultinomial.Group")(); g.counts.extend([0, 12, 70000, 3]); g.ratio = 41.625
    mine = wire.encode("dm.Group", {"counts": [0, 12, 70000, 3], "ratio": 41.625})
    assert mine == g.SerializeToString()
    back = wire.decode("dm.Group", g.SerializeToString())
    assert back["counts"] == [0, 12, 70000, 3] and back["ratio"] == 41.625
