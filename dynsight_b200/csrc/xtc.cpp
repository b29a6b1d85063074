// xtc.cpp -- GROMACS XTC trajectory ingest (host side, multi-threaded over frames).
//
// Replaces the MDAnalysis XTCReader behind Trj.init_from_xtc (trajectory/trajectory.py:68-79 of the
// reference).  The XTC container and its lossy integer coordinate
// compression are a published format (GROMACS "xdr3dfcoord", xdrfile library); this is an
// independent implementation of the decoder:
//
//   frame  := magic(1995) natoms step time box[9] natoms  (XDR: big endian, 4-byte units)
//             natoms <= 9 : 3*natoms floats
//             else        : precision minint[3] maxint[3] smallidx nbytes bytes[nbytes (+pad to 4)]
//   stream := bit stream, most significant bit first.  Every atom is either a "large" triple
//             (mixed-radix number in the bounding box sizes, or three fixed-width fields when a
//             size exceeds 24 bits) optionally followed by a run of "small" triples (mixed radix
//             in magic[smallidx]^3, offsets from the previous atom); the first small atom of a run
//             is swapped with its large predecessor (water oxygens / hydrogens); a 5-bit field
//             carries the run length and a -1/0/+1 adaptation of smallidx.
//
// Coordinates come back in angstrom (nm * 10 in float32, like MDAnalysis) so that the arrays are
// bit-identical to what the reference's Universe would hold.
#include <algorithm>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "../../include/dynsight_b200.h"

namespace dsg {
int set_error(int code, const char* fmt, ...);
}

namespace {

constexpr int kMagic = 1995;
constexpr int kFirstIdx = 9;
const int kMagicInts[] = {
    0,       0,       0,       0,       0,       0,        0,        0,        0,       8,
    10,      12,      16,      20,      25,      32,       40,       50,       64,      80,
    101,     128,     161,     203,     256,     322,      406,      512,      645,     812,
    1024,    1290,    1625,    2048,    2580,    3250,     4096,     5060,     6501,    8192,
    10321,   13003,   16384,   20642,   26007,   32768,    41285,    52015,    65536,   82570,
    104031,  131072,  165140,  208063,  262144,  330280,   416127,   524287,   660561,  832255,
    1048576, 1321122, 1664510, 2097152, 2642245, 3329021,  4194304,  5284491,  6658042, 8388607,
    10568983, 13316085, 16777216};
constexpr int kLastIdx = (int)(sizeof(kMagicInts) / sizeof(kMagicInts[0]));

inline uint32_t be32(const unsigned char* p) {
    return ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | (uint32_t)p[3];
}
inline float be_float(const unsigned char* p) {
    const uint32_t u = be32(p);
    float f;
    std::memcpy(&f, &u, 4);
    return f;
}

// bits needed for one value in [0, size]
int bits_for(unsigned size) {
    unsigned num = 1;
    int bits = 0;
    while (size >= num && bits < 32) {
        ++bits;
        num <<= 1;
    }
    return bits;
}

// bits needed for a mixed-radix number with the three given radices
int bits_for_product(const unsigned sizes[3]) {
    unsigned bytes[32];
    unsigned nbytes = 1;
    bytes[0] = 1;
    for (int i = 0; i < 3; ++i) {
        unsigned carry = 0, b = 0;
        for (; b < nbytes; ++b) {
            carry += bytes[b] * sizes[i];
            bytes[b] = carry & 0xffu;
            carry >>= 8;
        }
        while (carry != 0) {
            bytes[b++] = carry & 0xffu;
            carry >>= 8;
        }
        nbytes = b;
    }
    unsigned num = 1;
    int bits = 0;
    --nbytes;
    while (bytes[nbytes] >= num) {
        ++bits;
        num *= 2;
    }
    return bits + (int)nbytes * 8;
}

struct BitReader {
    const unsigned char* data;
    size_t size;
    size_t pos = 0;
    unsigned lastbits = 0;
    unsigned lastbyte = 0;
    bool overrun = false;

    unsigned next_byte() {
        if (pos >= size) {
            overrun = true;
            return 0;
        }
        return data[pos++];
    }
    int bits(int n) {
        const unsigned mask = n >= 32 ? 0xffffffffu : ((1u << n) - 1u);
        unsigned num = 0;
        while (n >= 8) {
            lastbyte = (lastbyte << 8) | next_byte();
            num |= (lastbyte >> lastbits) << (n - 8);
            n -= 8;
        }
        if (n > 0) {
            if (lastbits < (unsigned)n) {
                lastbits += 8;
                lastbyte = (lastbyte << 8) | next_byte();
            }
            lastbits -= (unsigned)n;
            num |= (lastbyte >> lastbits) & ((1u << n) - 1u);
        }
        return (int)(num & mask);
    }
    // three values packed as one mixed-radix number of `nbits` bits
    void triple(int nbits, const unsigned sizes[3], int out[3]) {
        int bytes[32];
        bytes[1] = bytes[2] = bytes[3] = 0;
        int nbytes = 0;
        while (nbits > 8) {
            bytes[nbytes++] = bits(8);
            nbits -= 8;
        }
        if (nbits > 0) bytes[nbytes++] = bits(nbits);
        for (int i = 2; i > 0; --i) {
            unsigned num = 0;
            for (int j = nbytes - 1; j >= 0; --j) {
                num = (num << 8) | (unsigned)bytes[j];
                const unsigned q = num / sizes[i];
                bytes[j] = (int)q;
                num -= q * sizes[i];
            }
            out[i] = (int)num;
        }
        out[0] = bytes[0] | (bytes[1] << 8) | (bytes[2] << 16) | (bytes[3] << 24);
    }
};

struct FrameHeader {
    int natoms = 0, step = 0;
    float time = 0.f;
    float box[9] = {0};
    size_t payload = 0;     // offset of the coordinate section relative to the frame start
    size_t frame_bytes = 0;  // full frame size
};

// Parses the fixed part of a frame at `p` (at least `avail` bytes readable).  Returns false when
// the bytes do not hold a complete frame.
bool parse_header(const unsigned char* p, size_t avail, FrameHeader* h) {
    if (avail < 56) return false;
    if ((int)be32(p) != kMagic) return false;
    h->natoms = (int)be32(p + 4);
    h->step = (int)be32(p + 8);
    h->time = be_float(p + 12);
    for (int i = 0; i < 9; ++i) h->box[i] = be_float(p + 16 + 4 * i);
    const int n2 = (int)be32(p + 52);
    if (n2 != h->natoms || h->natoms < 0) return false;
    h->payload = 56;
    if (h->natoms <= 9) {
        h->frame_bytes = 56 + (size_t)12 * h->natoms;
    } else {
        if (avail < 56 + 36) return false;
        const uint32_t nbytes = be32(p + 56 + 32);
        h->frame_bytes = 56 + 36 + (((size_t)nbytes + 3) & ~(size_t)3);
    }
    return h->frame_bytes <= avail;
}

// Decodes the coordinates of one frame into out[natoms*3] (angstrom).  Returns 0 or an error code.
int decode_frame(const unsigned char* p, const FrameHeader& h, float* out) {
    const int n = h.natoms;
    const float to_angstrom = 10.0f;
    const unsigned char* c = p + h.payload;
    if (n <= 9) {
        for (int i = 0; i < 3 * n; ++i) out[i] = be_float(c + 4 * i) * to_angstrom;
        return 0;
    }
    const float precision = be_float(c);
    int minint[3], maxint[3];
    for (int k = 0; k < 3; ++k) minint[k] = (int)be32(c + 4 + 4 * k);
    for (int k = 0; k < 3; ++k) maxint[k] = (int)be32(c + 16 + 4 * k);
    int smallidx = (int)be32(c + 28);
    const uint32_t nbytes = be32(c + 32);
    if (smallidx < kFirstIdx || smallidx >= kLastIdx || !(precision > 0.f)) return 1;
    unsigned sizeint[3];
    int bitsizeint[3] = {0, 0, 0};
    for (int k = 0; k < 3; ++k) sizeint[k] = (unsigned)(maxint[k] - minint[k] + 1);
    int bitsize;
    if ((sizeint[0] | sizeint[1] | sizeint[2]) > 0xffffffu) {
        for (int k = 0; k < 3; ++k) bitsizeint[k] = bits_for(sizeint[k]);
        bitsize = 0;  // three fixed-width fields
    } else {
        bitsize = bits_for_product(sizeint);
    }
    int smaller = kMagicInts[std::max(kFirstIdx, smallidx - 1)] / 2;
    int smallnum = kMagicInts[smallidx] / 2;
    unsigned sizesmall[3] = {(unsigned)kMagicInts[smallidx], (unsigned)kMagicInts[smallidx],
                             (unsigned)kMagicInts[smallidx]};
    const float inv_precision = (float)(1.0 / precision);
    BitReader br{c + 36, nbytes};
    int run = 0, i = 0;
    float* w = out;
    int cur[3], prev[3];
    auto emit = [&](const int v[3]) {
        w[0] = (float)v[0] * inv_precision * to_angstrom;
        w[1] = (float)v[1] * inv_precision * to_angstrom;
        w[2] = (float)v[2] * inv_precision * to_angstrom;
        w += 3;
    };
    while (i < n) {
        if (bitsize == 0) {
            cur[0] = br.bits(bitsizeint[0]);
            cur[1] = br.bits(bitsizeint[1]);
            cur[2] = br.bits(bitsizeint[2]);
        } else {
            br.triple(bitsize, sizeint, cur);
        }
        ++i;
        for (int k = 0; k < 3; ++k) {
            cur[k] += minint[k];
            prev[k] = cur[k];
        }
        const int flag = br.bits(1);
        int is_smaller = 0;
        if (flag == 1) {
            run = br.bits(5);
            is_smaller = run % 3;
            run -= is_smaller;
            --is_smaller;
        }
        if (run > 0) {
            if (i + run / 3 > n) return 2;
            for (int k = 0; k < run; k += 3) {
                br.triple(smallidx, sizesmall, cur);
                ++i;
                for (int d = 0; d < 3; ++d) cur[d] += prev[d] - smallnum;
                if (k == 0) {
                    // the first small atom changes places with its large predecessor
                    for (int d = 0; d < 3; ++d) std::swap(cur[d], prev[d]);
                    emit(prev);
                } else {
                    for (int d = 0; d < 3; ++d) prev[d] = cur[d];
                }
                emit(cur);
            }
        } else {
            emit(cur);
        }
        smallidx += is_smaller;
        if (smallidx < kFirstIdx || smallidx >= kLastIdx) return 3;
        if (is_smaller < 0) {
            smallnum = smaller;
            smaller = smallidx > kFirstIdx ? kMagicInts[smallidx - 1] / 2 : 0;
        } else if (is_smaller > 0) {
            smaller = smallnum;
            smallnum = kMagicInts[smallidx] / 2;
        }
        sizesmall[0] = sizesmall[1] = sizesmall[2] = (unsigned)kMagicInts[smallidx];
        if (br.overrun) return 4;
    }
    return 0;
}

bool read_file(const char* path, std::vector<unsigned char>* data) {
    FILE* fh = std::fopen(path, "rb");
    if (!fh) return false;
    std::fseek(fh, 0, SEEK_END);
    const long sz = std::ftell(fh);
    std::fseek(fh, 0, SEEK_SET);
    if (sz < 0) {
        std::fclose(fh);
        return false;
    }
    data->resize((size_t)sz);
    const size_t got = sz > 0 ? std::fread(data->data(), 1, (size_t)sz, fh) : 0;
    std::fclose(fh);
    return got == (size_t)sz;
}

}  // namespace

using dsg::set_error;

extern "C" int dsg_xtc_index(const char* path, int64_t* n_frames_out, int32_t* n_atoms_out,
                             int64_t* offsets_out, int64_t max_frames) {
    if (!path || !n_frames_out || !n_atoms_out)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "NULL argument");
    std::vector<unsigned char> data;
    if (!read_file(path, &data)) return set_error(DSG_ERR_INVALID_ARGUMENT, "cannot read %s", path);
    size_t off = 0;
    int64_t nf = 0;
    int natoms = -1;
    while (off < data.size()) {
        FrameHeader h;
        if (!parse_header(data.data() + off, data.size() - off, &h))
            return set_error(DSG_ERR_INVALID_ARGUMENT, "%s: bad or truncated XTC frame %lld at byte %zu",
                             path, (long long)nf, off);
        if (natoms < 0) natoms = h.natoms;
        if (h.natoms != natoms)
            return set_error(DSG_ERR_INVALID_ARGUMENT, "%s: atom count changes at frame %lld", path,
                             (long long)nf);
        if (offsets_out && nf < max_frames) offsets_out[nf] = (int64_t)off;
        ++nf;
        off += h.frame_bytes;
    }
    *n_frames_out = nf;
    *n_atoms_out = natoms < 0 ? 0 : natoms;
    return DSG_OK;
}

extern "C" int dsg_xtc_read(const char* path, const int64_t* offsets, int64_t n_frames,
                            int32_t n_atoms, float* pos_out, float* box_out, int32_t* step_out,
                            float* time_out, int n_threads) {
    if (!path || !offsets || !pos_out || n_frames < 0 || n_atoms < 0)
        return set_error(DSG_ERR_INVALID_ARGUMENT, "NULL or negative argument");
    std::vector<unsigned char> data;
    if (!read_file(path, &data)) return set_error(DSG_ERR_INVALID_ARGUMENT, "cannot read %s", path);
    if (n_threads < 1) n_threads = (int)std::max(1u, std::thread::hardware_concurrency());
    n_threads = (int)std::min<int64_t>(n_threads, std::max<int64_t>(1, n_frames));
    std::vector<int> status((size_t)n_threads, 0);
    std::vector<int64_t> bad((size_t)n_threads, -1);
    auto work = [&](int t) {
        for (int64_t f = t; f < n_frames; f += n_threads) {
            const size_t off = (size_t)offsets[f];
            FrameHeader h;
            if (off >= data.size() || !parse_header(data.data() + off, data.size() - off, &h) ||
                h.natoms != n_atoms) {
                status[t] = 10;
                bad[t] = f;
                return;
            }
            const int rc = decode_frame(data.data() + off, h, pos_out + (size_t)f * n_atoms * 3);
            if (rc) {
                status[t] = rc;
                bad[t] = f;
                return;
            }
            if (box_out)
                for (int i = 0; i < 9; ++i) box_out[f * 9 + i] = h.box[i] * 10.0f;
            if (step_out) step_out[f] = h.step;
            if (time_out) time_out[f] = h.time;
        }
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < n_threads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    for (int t = 0; t < n_threads; ++t)
        if (status[t])
            return set_error(DSG_ERR_INVALID_ARGUMENT, "%s: frame %lld does not decode (code %d)",
                             path, (long long)bad[t], status[t]);
    return DSG_OK;
}
