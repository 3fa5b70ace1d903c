// obs_codec.cu -- host side of the observation wire format: the data format between
// `varlociraptor preprocess` (part 1) and `call` (part 2).  Replaces read_observations /
// write_observations (reference src/calling/variants/preprocessing/mod.rs:452-598,
// OBSERVATION_FORMAT_VERSION "8") and MiniLogProb (src/utils/mod.rs:381-407), producing /
// consuming the struct-of-arrays columns the likelihood kernels read (include/vlr_gpu.h).
// Host-only code (bincode byte shuffling); the layouts are pinned against records written by the
// reference itself (tests/golden/obs_records.json).
#include <math.h>
#include <string.h>

#include <vector>

#include "../../include/vlr_gpu.h"

namespace {

// half::f16::from_f64: round to nearest even, subnormals kept, overflow to infinity
uint16_t f64_to_f16_bits(double d) {
    uint64_t x;
    memcpy(&x, &d, 8);
    const uint16_t sign = (uint16_t)((x >> 48) & 0x8000u);
    const int exp = (int)((x >> 52) & 0x7ff);
    const uint64_t man = x & 0x000fffffffffffffull;
    if (exp == 0x7ff) return (uint16_t)(sign | 0x7c00u | (man ? 0x0200u | (uint16_t)(man >> 42) : 0));
    const int e = exp - 1023;  // unbiased
    if (e > 15) return (uint16_t)(sign | 0x7c00u);
    if (e >= -14) {  // normal half
        uint32_t half_man = (uint32_t)(man >> 42);
        const uint64_t rest = man & ((1ull << 42) - 1), halfway = 1ull << 41;
        uint32_t h = ((uint32_t)(e + 15) << 10) | half_man;
        if (rest > halfway || (rest == halfway && (half_man & 1))) h++;  // may carry into the exponent
        return (uint16_t)(sign | h);
    }
    if (e < -25) return sign;  // below half of the smallest subnormal
    // subnormal half: value = m * 2^-24
    const uint64_t full = man | (1ull << 52);       // 1.man
    const int shift = 42 + (-14 - e);               // bits dropped
    const uint32_t half_man = (uint32_t)(full >> shift);
    const uint64_t rest = full & ((1ull << shift) - 1), halfway = 1ull << (shift - 1);
    uint32_t h = half_man;
    if (rest > halfway || (rest == halfway && (half_man & 1))) h++;
    return (uint16_t)(sign | h);
}
double f16_bits_to_f64(uint16_t h) {
    const int sign = h >> 15, exp = (h >> 10) & 31, man = h & 1023;
    double v;
    if (exp == 31) v = man ? NAN : INFINITY;
    else if (exp == 0) v = ldexp((double)man, -24);
    else v = ldexp((double)(man | 1024), exp - 25);
    return sign ? -v : v;
}

double ln_one_minus_exp(double p) { return p < -0.693 ? log1p(-exp(p)) : log(-expm1(p)); }

struct Reader {
    std::vector<uint8_t> b;
    size_t pos = 0;
    bool ok = true;
    Reader(const int32_t *v, int32_t n) {
        b.resize((size_t)n * 2);
        for (int32_t k = 0; k < n; ++k) {  // i32 -> u16 -> two bytes, little endian (mod.rs:473-479)
            b[2 * k] = (uint8_t)(v[k] & 0xff);
            b[2 * k + 1] = (uint8_t)((v[k] >> 8) & 0xff);
        }
    }
    template <typename T>
    T get() {
        T t = 0;
        if (pos + sizeof(T) > b.size()) { ok = false; return t; }
        memcpy(&t, &b[pos], sizeof(T));
        pos += sizeof(T);
        return t;
    }
    bool at_end() const { return ok && (pos == b.size() || (pos + 1 == b.size() && b[pos] == 0)); }
};

struct Writer {
    std::vector<uint8_t> b;
    template <typename T>
    void put(T t) {
        const size_t p = b.size();
        b.resize(p + sizeof(T));
        memcpy(&b[p], &t, sizeof(T));
    }
    // pad to even, u16 little endian as i32 (mod.rs:559-575); returns the number of values or -1
    int64_t emit(int32_t *out, int64_t cap) {
        if (b.size() & 1) b.push_back(0);
        const int64_t n = (int64_t)b.size() / 2;
        if (n > cap) return -1;
        for (int64_t k = 0; k < n; ++k) out[k] = (int32_t)(b[2 * k] | (b[2 * k + 1] << 8));
        return n;
    }
};

}  // namespace

extern "C" {

int64_t vlr_obs_encoded_bound(int64_t n_obs) {
    // largest field: u64 length + n * (u32 variant + f32) bytes, padded, two bytes per value
    return (8 + n_obs * 8 + 1) / 2 + 1;
}

int vlr_obs_decode(const int32_t *const *fields, const int32_t *n_values, int64_t capacity,
                   int64_t *out_n_obs, double *const *out_cols, uint16_t *out_cat) {
    if (!fields || !n_values || !out_n_obs) return VLR_ERR_INVALID;
    int64_t n = -1;
    // ---- the seven MiniLogProb vectors (utils/mod.rs:381-407)
    static const int col_of_field[7] = {0, 3, 2, 4, 5, 6, 8};  // field order -> vlr_obs_columns order
    for (int f = 0; f < 7; ++f) {
        if (!fields[f] || n_values[f] < 4) return VLR_ERR_INVALID;
        Reader r(fields[f], n_values[f]);
        const uint64_t len = r.get<uint64_t>();
        if (n < 0) {
            n = (int64_t)len;
            *out_n_obs = n;
            if (!out_cols && !out_cat) return VLR_OK;  // size query
            if (n > capacity || !out_cols || !out_cat) return VLR_ERR_INVALID;
        }
        if ((int64_t)len != n) return VLR_ERR_INVALID;
        double *dst = out_cols[col_of_field[f]];
        if (!dst) return VLR_ERR_INVALID;
        for (int64_t k = 0; k < n; ++k) {
            const uint32_t var = r.get<uint32_t>();
            if (var == 0) dst[k] = f16_bits_to_f64(r.get<uint16_t>());
            else if (var == 1) dst[k] = (double)r.get<float>();
            else return VLR_ERR_INVALID;
        }
        if (!r.at_end()) return VLR_ERR_INVALID;
    }
    // derived columns (ObservationBuilder, observation.rs:218-226)
    if (!out_cols[1] || !out_cols[7]) return VLR_ERR_INVALID;
    for (int64_t k = 0; k < n; ++k) {
        out_cols[1][k] = ln_one_minus_exp(out_cols[0][k]);
        out_cols[7][k] = ln_one_minus_exp(out_cols[6][k]);
    }
    for (int64_t k = 0; k < n; ++k) out_cat[k] = 0;
    // ---- enum vectors: u32 variant indices
    static const int enum_field[4] = {VLR_OBS_STRAND, VLR_OBS_READ_ORIENTATION, VLR_OBS_READ_POSITION,
                                      VLR_OBS_INDEL_OPERATIONS};
    static const int enum_shift[4] = {0, 2, 6, 8};
    static const uint32_t enum_n[4] = {4, 9, 2, 3};
    for (int e = 0; e < 4; ++e) {
        const int f = enum_field[e];
        if (!fields[f] || n_values[f] == 0) {
            if (f != VLR_OBS_INDEL_OPERATIONS) return VLR_ERR_INVALID;
            for (int64_t k = 0; k < n; ++k) out_cat[k] |= (uint16_t)(2u << 8);  // format 7: IndelOperations::None
            continue;
        }
        Reader r(fields[f], n_values[f]);
        if ((int64_t)r.get<uint64_t>() != n) return VLR_ERR_INVALID;
        for (int64_t k = 0; k < n; ++k) {
            const uint32_t v = r.get<uint32_t>();
            if (v >= enum_n[e]) return VLR_ERR_INVALID;
            out_cat[k] |= (uint16_t)(v << enum_shift[e]);
        }
        if (!r.at_end()) return VLR_ERR_INVALID;
    }
    // ---- BitVec<u8>: Option tag, u64 block count, blocks, u64 bit count
    static const int bit_field[2] = {VLR_OBS_SOFTCLIPPED, VLR_OBS_PAIRED};
    static const int bit_shift[2] = {7, 10};
    for (int e = 0; e < 2; ++e) {
        const int f = bit_field[e];
        if (!fields[f] || n_values[f] < 4) return VLR_ERR_INVALID;
        Reader r(fields[f], n_values[f]);
        const uint8_t tag = r.get<uint8_t>();
        if (tag == 0) {
            if (r.get<uint64_t>() != 0 || n != 0) return VLR_ERR_INVALID;
        } else if (tag == 1) {
            const uint64_t nb = r.get<uint64_t>();
            if ((int64_t)nb != (n + 7) / 8) return VLR_ERR_INVALID;
            const size_t p0 = r.pos;
            r.pos += nb;
            if (r.pos > r.b.size() || (int64_t)r.get<uint64_t>() != n) return VLR_ERR_INVALID;
            for (int64_t k = 0; k < n; ++k)
                if ((r.b[p0 + (size_t)(k >> 3)] >> (k & 7)) & 1) out_cat[k] |= (uint16_t)(1u << bit_shift[e]);
        } else {
            return VLR_ERR_INVALID;
        }
        if (!r.at_end()) return VLR_ERR_INVALID;
    }
    return VLR_OK;
}

int vlr_obs_encode(int64_t n_obs, const double *const *cols, const uint16_t *cat,
                   int32_t *const *out_fields, int64_t capacity, int32_t *out_n_values) {
    if (n_obs < 0 || !cols || (!cat && n_obs) || !out_fields || !out_n_values) return VLR_ERR_INVALID;
    static const int col_of_field[7] = {0, 3, 2, 4, 5, 6, 8};
    for (int f = 0; f < 7; ++f) {
        const double *src = cols[col_of_field[f]];
        if (!src && n_obs) return VLR_ERR_INVALID;
        Writer w;
        w.put<uint64_t>((uint64_t)n_obs);
        for (int64_t k = 0; k < n_obs; ++k) {
            // MiniLogProb::new: f16 iff p < -10 and the f16 value has the same integer part
            const double p = src[k];
            const uint16_t h = f64_to_f16_bits(p);
            const double proj = f16_bits_to_f64(h);
            if (p < -10.0 && (long long)floor(proj) == (long long)floor(p)) {
                w.put<uint32_t>(0);
                w.put<uint16_t>(h);
            } else {
                w.put<uint32_t>(1);
                w.put<float>((float)p);
            }
        }
        const int64_t nv = w.emit(out_fields[f], capacity);
        if (nv < 0) return VLR_ERR_INVALID;
        out_n_values[f] = (int32_t)nv;
    }
    static const int enum_field[4] = {VLR_OBS_STRAND, VLR_OBS_READ_ORIENTATION, VLR_OBS_READ_POSITION,
                                      VLR_OBS_INDEL_OPERATIONS};
    static const int enum_shift[4] = {0, 2, 6, 8};
    static const uint32_t enum_mask[4] = {3, 15, 1, 3};
    for (int e = 0; e < 4; ++e) {
        Writer w;
        w.put<uint64_t>((uint64_t)n_obs);
        for (int64_t k = 0; k < n_obs; ++k) w.put<uint32_t>((cat[k] >> enum_shift[e]) & enum_mask[e]);
        const int64_t nv = w.emit(out_fields[enum_field[e]], capacity);
        if (nv < 0) return VLR_ERR_INVALID;
        out_n_values[enum_field[e]] = (int32_t)nv;
    }
    static const int bit_field[2] = {VLR_OBS_SOFTCLIPPED, VLR_OBS_PAIRED};
    static const int bit_shift[2] = {7, 10};
    for (int e = 0; e < 2; ++e) {
        Writer w;
        if (n_obs == 0) {
            w.put<uint8_t>(0);
            w.put<uint64_t>(0);
        } else {
            const int64_t nb = (n_obs + 7) / 8;
            w.put<uint8_t>(1);
            w.put<uint64_t>((uint64_t)nb);
            std::vector<uint8_t> blocks((size_t)nb, 0);
            for (int64_t k = 0; k < n_obs; ++k)
                if ((cat[k] >> bit_shift[e]) & 1) blocks[(size_t)(k >> 3)] |= (uint8_t)(1u << (k & 7));
            for (uint8_t b : blocks) w.put<uint8_t>(b);
            w.put<uint64_t>((uint64_t)n_obs);
        }
        const int64_t nv = w.emit(out_fields[bit_field[e]], capacity);
        if (nv < 0) return VLR_ERR_INVALID;
        out_n_values[bit_field[e]] = (int32_t)nv;
    }
    return VLR_OK;
}

int vlr_obs_encode_batch(int64_t n_records, const double *const *cols, const uint16_t *cat, const int64_t *obs_off,
                         uint16_t *const *out_words, int64_t *const *out_word_off, int64_t capacity_words) {
    if (n_records < 0 || !cols || !cat || !obs_off || !out_words || !out_word_off) return VLR_ERR_INVALID;
    for (int f = 0; f < VLR_OBS_N_FIELDS; ++f) {
        if (!out_words[f] || !out_word_off[f]) return VLR_ERR_INVALID;
        out_word_off[f][0] = 0;
    }
    std::vector<int32_t> tmp[VLR_OBS_N_FIELDS];
    for (int64_t r = 0; r < n_records; ++r) {
        const int64_t r0 = obs_off[r], n = obs_off[r + 1] - r0;
        if (n < 0) return VLR_ERR_INVALID;
        const int64_t cap = vlr_obs_encoded_bound(n);
        const double *rc[9];
        int32_t *fp[VLR_OBS_N_FIELDS];
        int32_t nv[VLR_OBS_N_FIELDS];
        for (int c = 0; c < 9; ++c) rc[c] = cols[c] ? cols[c] + r0 : nullptr;
        for (int f = 0; f < VLR_OBS_N_FIELDS; ++f) {
            if ((int64_t)tmp[f].size() < cap) tmp[f].resize((size_t)cap);
            fp[f] = tmp[f].data();
        }
        const int st = vlr_obs_encode(n, rc, cat + r0, fp, cap, nv);
        if (st != VLR_OK) return st;
        for (int f = 0; f < VLR_OBS_N_FIELDS; ++f) {
            const int64_t w0 = out_word_off[f][r];
            if (w0 + nv[f] > capacity_words) return VLR_ERR_INVALID;
            for (int32_t k = 0; k < nv[f]; ++k) out_words[f][w0 + k] = (uint16_t)fp[f][k];
            out_word_off[f][r + 1] = w0 + nv[f];
        }
    }
    return VLR_OK;
}

}  // extern "C"
