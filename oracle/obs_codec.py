"""CPU restatement (numpy) of Varlociraptor's observation wire format -- TEST INFRASTRUCTURE ONLY.

Follows src/calling/variants/preprocessing/mod.rs:452-598 (read_observations / write_observations,
OBSERVATION_FORMAT_VERSION "8") and src/utils/mod.rs:381-407 (MiniLogProb).  Each observation field
is a bincode 1.3 serialisation, padded to even length, cut into little-endian u16 and stored as a
BCF INFO integer vector (mod.rs:559-581).

bincode layouts used here (third-party crates not in /root/reference: bincode 1.3.3, half 1.7.1,
bv 0.11.1, bio-types 0.12.0) are PINNED against the observation records embedded in the
reference's own testcases (tests/resources/testcases/*/candidates.vcf, format version 7 = version 8
without INDEL_OPERATIONS; extracted by tests/golden/make_obs_golden.py): every field must be
consumed exactly (up to the one pad byte) and all fields of a record must agree on the number of
observations.
  Vec<T>              u64 LE length, then the elements
  MiniLogProb         u32 LE variant index (0 = F16, 1 = F32), then the f16 (2 bytes) / f32 (4 bytes)
  Strand, ReadPosition, IndelOperations, SequenceReadPairOrientation
                      u32 LE variant index (observation.rs:40-45, 118-121, 130-134; bio-types:
                      F1R2 F2R1 R1F2 R2F1 F1F2 R1R2 F2F1 R2R1 None)
  BitVec<u8>          Option tag byte (1 = Some), u64 LE number of blocks, the blocks (bit i of the
                      vector = bit i%8 of block i/8), u64 LE number of bits; an empty vector is
                      the tag byte 0 followed by the bit count
"""
import struct

import numpy as np

PROB_FIELDS = ["PROB_MAPPING", "PROB_REF", "PROB_ALT", "PROB_MISSED_ALLELE", "PROB_SAMPLE_ALT",
               "PROB_DOUBLE_OVERLAP", "PROB_HIT_BASE"]
ENUM_FIELDS = ["STRAND", "READ_ORIENTATION", "READ_POSITION", "INDEL_OPERATIONS"]
BIT_FIELDS = ["SOFTCLIPPED", "PAIRED"]
# order of the library's field arrays (include/vlr_gpu.h, VLR_OBS_*)
FIELDS = ["PROB_MAPPING", "PROB_REF", "PROB_ALT", "PROB_MISSED_ALLELE", "PROB_SAMPLE_ALT",
          "PROB_DOUBLE_OVERLAP", "PROB_HIT_BASE", "STRAND", "READ_ORIENTATION", "READ_POSITION",
          "SOFTCLIPPED", "INDEL_OPERATIONS", "PAIRED"]


def ints_to_bytes(ints):
    """i32 INFO values -> u16 LE -> bytes (mod.rs:473-479)."""
    return b"".join(struct.pack("<H", int(v) & 0xFFFF) for v in ints)


def bytes_to_ints(b):
    """bytes -> pad to even -> u16 LE as i32 (mod.rs:559-575)."""
    if len(b) % 2:
        b += b"\0"
    return [struct.unpack_from("<H", b, k)[0] for k in range(0, len(b), 2)]


def minilogprob(p):
    """MiniLogProb::new + to_logprob (utils/mod.rs:391-406): (variant, de-quantised value)."""
    p = float(p)
    with np.errstate(over="ignore"):
        half = np.float16(p)
    proj = float(half)
    if p < -10.0 and np.isfinite(proj) and np.floor(proj) == np.floor(p):
        return 0, proj
    if p < -10.0 and not np.isfinite(p) and proj == p:  # -inf: floor(-inf) == floor(-inf)
        return 0, proj
    with np.errstate(over="ignore"):
        return 1, float(np.float32(p))


def _check_end(b, pos, tag):
    rest = b[pos:]
    assert rest in (b"", b"\0"), "%s: %d undecoded bytes" % (tag, len(rest))


def decode_probs(ints, tag="PROB"):
    b = ints_to_bytes(ints)
    n = struct.unpack_from("<Q", b, 0)[0]
    pos, out = 8, np.empty(n, np.float64)
    for k in range(n):
        var = struct.unpack_from("<I", b, pos)[0]
        pos += 4
        if var == 0:
            out[k] = float(np.frombuffer(b, np.float16, 1, pos)[0])
            pos += 2
        elif var == 1:
            out[k] = float(struct.unpack_from("<f", b, pos)[0])
            pos += 4
        else:
            raise ValueError("%s: MiniLogProb variant %d" % (tag, var))
    _check_end(b, pos, tag)
    return out


def decode_enums(ints, n_variants, tag="ENUM"):
    b = ints_to_bytes(ints)
    n = struct.unpack_from("<Q", b, 0)[0]
    out = np.frombuffer(b, "<u4", n, 8).astype(np.int32)
    assert (out < n_variants).all(), tag
    _check_end(b, 8 + 4 * n, tag)
    return out


def decode_bits(ints, tag="BITS"):
    b = ints_to_bytes(ints)
    if b[0] == 0:
        n = struct.unpack_from("<Q", b, 1)[0]
        assert n == 0, tag
        _check_end(b, 9, tag)
        return np.zeros(0, bool)
    assert b[0] == 1, tag
    nblocks = struct.unpack_from("<Q", b, 1)[0]
    blocks = np.frombuffer(b, np.uint8, nblocks, 9)
    n = struct.unpack_from("<Q", b, 9 + nblocks)[0]
    assert nblocks == (n + 7) // 8, tag
    _check_end(b, 17 + nblocks, tag)
    return np.unpackbits(blocks, bitorder="little")[:n].astype(bool)


def ln_one_minus_exp(p):
    """LogProb::ln_one_minus_exp (bio 0.37)."""
    p = np.asarray(p, np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        return np.where(p < -0.693, np.log1p(-np.exp(p)), np.log(-np.expm1(p)))


def decode_record(fields):
    """fields: dict tag -> list of INFO ints.  Returns the pileup as the nine fp64 columns of
    include/vlr_gpu.h plus the packed categoricals (read_observations, mod.rs:455-527;
    prob_mismapping / prob_single_overlap as ObservationBuilder derives them,
    observation.rs:218-226).  A record without INDEL_OPERATIONS (format 7) gets
    IndelOperations::None."""
    cols = {}
    for tag, name in zip(PROB_FIELDS, ["prob_mapping", "prob_ref", "prob_alt", "prob_missed_allele",
                                       "prob_sample_alt", "prob_double_overlap", "prob_hit_base"]):
        cols[name] = decode_probs(fields[tag], tag)
    n = len(cols["prob_mapping"])
    strand = decode_enums(fields["STRAND"], 4, "STRAND")
    orient = decode_enums(fields["READ_ORIENTATION"], 9, "READ_ORIENTATION")
    rpos = decode_enums(fields["READ_POSITION"], 2, "READ_POSITION")
    indel = (decode_enums(fields["INDEL_OPERATIONS"], 3, "INDEL_OPERATIONS")
             if fields.get("INDEL_OPERATIONS") else np.full(n, 2, np.int32))
    soft = decode_bits(fields["SOFTCLIPPED"], "SOFTCLIPPED")
    paired = decode_bits(fields["PAIRED"], "PAIRED")
    for a in list(cols.values()) + [strand, orient, rpos, indel, soft, paired]:
        assert len(a) == n, "fields disagree on the number of observations"
    cols["prob_mismapping"] = ln_one_minus_exp(cols["prob_mapping"])
    cols["prob_single_overlap"] = ln_one_minus_exp(cols["prob_double_overlap"])
    cat = (strand | (orient << 2) | (rpos << 6) | (soft.astype(np.int32) << 7) | (indel << 8) |
           (paired.astype(np.int32) << 10)).astype(np.uint16)
    return cols, cat


def encode_probs(values):
    out = [struct.pack("<Q", len(values))]
    for p in values:
        var, q = minilogprob(p)
        out.append(struct.pack("<I", var))
        out.append(np.float16(q).tobytes() if var == 0 else struct.pack("<f", q))
    return bytes_to_ints(b"".join(out))


def encode_enums(values):
    return bytes_to_ints(struct.pack("<Q", len(values)) + np.asarray(values, "<u4").tobytes())


def encode_bits(bits):
    bits = np.asarray(bits, bool)
    if len(bits) == 0:
        return bytes_to_ints(b"\0" + struct.pack("<Q", 0))
    blocks = np.packbits(bits, bitorder="little")
    return bytes_to_ints(b"\1" + struct.pack("<Q", len(blocks)) + blocks.tobytes() +
                         struct.pack("<Q", len(bits)))


def encode_record(cols, cat):
    """write_observations (mod.rs:529-598): dict tag -> list of INFO ints (format 8)."""
    cat = np.asarray(cat, np.int32)
    f = {}
    for tag, name in zip(PROB_FIELDS, ["prob_mapping", "prob_ref", "prob_alt", "prob_missed_allele",
                                       "prob_sample_alt", "prob_double_overlap", "prob_hit_base"]):
        f[tag] = encode_probs(cols[name])
    f["STRAND"] = encode_enums(cat & 3)
    f["READ_ORIENTATION"] = encode_enums((cat >> 2) & 15)
    f["READ_POSITION"] = encode_enums((cat >> 6) & 1)
    f["SOFTCLIPPED"] = encode_bits((cat >> 7) & 1)
    f["INDEL_OPERATIONS"] = encode_enums((cat >> 8) & 3)
    f["PAIRED"] = encode_bits((cat >> 10) & 1)
    return f
