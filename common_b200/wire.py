"""Protobuf wire format of the reference's checkpoint messages, hand-rolled (no protoc in this image).

In-tree messages (microscopes/io/schema.proto:3-55): CRP, GroupData, GroupManager, MixtureModelGroup,
MixtureModelState -- what group_manager::serialize (group_manager.hpp:285-298) and the mixture-model state
write.  The per-family Shared / Group messages live in the un-vendored `distributions` library
(distributions/io/schema.proto, reached through distributions.hpp:300-314,355-369); their field numbers below
are restated from that library's published schema and are UNPINNED here (nothing in the reference tree holds
them).  Floats are 32-bit on the wire, like the reference's members.

Only what the messages above need is implemented: varint (wire type 0), 32-bit (5), length-delimited (2),
packed and unpacked repeated scalars on input, unpacked on output (proto2 default).
"""
import struct

VARINT, FIXED64, BYTES, FIXED32 = 0, 1, 2, 5

# field spec: number -> (name, kind, repeated); kind in {"float", "uint", "int", "bytes", message-name}
SCHEMA = {
    # ---- microscopes/io/schema.proto ----
    "CRP": {1: ("alpha", "float", False)},
    "GroupData": {1: ("id", "uint", False), 2: ("data", "bytes", False)},
    "GroupManager": {1: ("alpha", "float", False), 2: ("assignments", "int", True), 3: ("groups", "GroupData", True)},
    "MixtureModelGroup": {1: ("suffstats", "bytes", True)},
    "MixtureModelState": {1: ("hypers", "bytes", True), 2: ("groups", "bytes", False)},
    "bbnc.Shared": {1: ("alpha", "float", False), 2: ("beta", "float", False)},                 # schema.proto:8-12
    "bbnc.Group": {1: ("p", "float", False), 2: ("heads", "uint", False), 3: ("tails", "uint", False)},  # schema.proto:14-18
    "dm.Shared": {1: ("alphas", "float", True)},                                                # schema.proto:21-23
    "dm.Group": {1: ("counts", "uint", True), 2: ("ratio", "float", False)},                    # schema.proto:25-28
    # ---- distributions/io/schema.proto [R: unpinned] ----
    "bb.Shared": {1: ("alpha", "float", False), 2: ("beta", "float", False)},
    "bb.Group": {1: ("heads", "uint", False), 2: ("tails", "uint", False)},
    "bnb.Shared": {1: ("alpha", "float", False), 2: ("beta", "float", False), 3: ("r", "uint", False)},
    "bnb.Group": {1: ("count", "uint", False), 2: ("sum", "uint", False)},
    "gp.Shared": {1: ("alpha", "float", False), 2: ("inv_beta", "float", False)},
    "gp.Group": {1: ("count", "uint", False), 2: ("sum", "uint", False), 3: ("log_prod", "float", False)},
    "nich.Shared": {1: ("mu", "float", False), 2: ("kappa", "float", False), 3: ("sigmasq", "float", False), 4: ("nu", "float", False)},
    "nich.Group": {1: ("count", "uint", False), 2: ("mean", "float", False), 3: ("count_times_variance", "float", False)},
    "dd.Shared": {1: ("alphas", "float", True)},
    "dd.Group": {1: ("counts", "uint", True)},
    "niw.Shared": {1: ("mu", "float", True), 2: ("kappa", "float", False), 3: ("psi", "float", True), 4: ("nu", "float", False)},
    "niw.Group": {1: ("count", "uint", False), 2: ("sum_x", "float", True), 3: ("sum_xxT", "float", True)},
}


def _varint(v):
    v &= (1 << 64) - 1          # negative int32 / int64 are sign-extended to 64 bits (10 bytes), as protobuf does
    out = bytearray()
    while True:
        b = v & 0x7F
        v >>= 7
        if v:
            out.append(b | 0x80)
        else:
            out.append(b)
            return bytes(out)


def _read_varint(buf, pos):
    shift = v = 0
    while True:
        b = buf[pos]
        pos += 1
        v |= (b & 0x7F) << shift
        if not b & 0x80:
            return v, pos
        shift += 7
        if shift > 70:
            raise ValueError("malformed varint")


def encode(msg, values):
    """values: dict name -> value (list for repeated fields); fields absent from the dict are not written"""
    spec = SCHEMA[msg]
    out = bytearray()
    for num in sorted(spec):
        name, kind, rep = spec[num]
        if name not in values:
            continue
        items = values[name] if rep else [values[name]]
        for v in items:
            if kind == "float":
                out += _varint((num << 3) | FIXED32) + struct.pack("<f", float(v))
            elif kind in ("uint", "int"):
                out += _varint((num << 3) | VARINT) + _varint(int(v))
            else:
                payload = v if kind == "bytes" else encode(kind, v)
                out += _varint((num << 3) | BYTES) + _varint(len(payload)) + bytes(payload)
    return bytes(out)


def decode(msg, buf):
    spec = SCHEMA[msg]
    buf = bytes(buf)
    out = {name: [] for name, _, rep in spec.values() if rep}
    pos = 0
    while pos < len(buf):
        key, pos = _read_varint(buf, pos)
        num, wt = key >> 3, key & 7
        if wt == VARINT:
            raw, pos = _read_varint(buf, pos)
        elif wt == FIXED32:
            raw = buf[pos:pos + 4]; pos += 4
        elif wt == FIXED64:
            raw = buf[pos:pos + 8]; pos += 8
        elif wt == BYTES:
            n, pos = _read_varint(buf, pos)
            raw = buf[pos:pos + n]; pos += n
        else:
            raise ValueError("unsupported wire type %d" % wt)
        if num not in spec:
            continue                                  # unknown field: skipped, as protobuf does
        name, kind, rep = spec[num]
        if kind == "float":
            vals = ([struct.unpack("<f", raw)[0]] if wt == FIXED32 else
                    [x[0] for x in struct.iter_unpack("<f", raw)])            # packed
        elif kind in ("uint", "int"):
            if wt == VARINT:
                vals = [raw]
            else:                                                             # packed
                vals, p2 = [], 0
                while p2 < len(raw):
                    x, p2 = _read_varint(raw, p2)
                    vals.append(x)
            if kind == "int":
                vals = [x - (1 << 64) if x >= (1 << 63) else x for x in vals]
        elif kind == "bytes":
            vals = [raw]
        else:
            vals = [decode(kind, raw)]
        if rep:
            out[name].extend(vals)
        else:
            out[name] = vals[-1]
    return out
