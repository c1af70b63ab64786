// emu.cpp -- sequential emulation of the CUDA kernels' tile decomposition on the
// CPU, calling the SAME per-thread functions (shafa-cd_b200/csrc/core.h) the
// kernels call, checked against the oracle (oracle/liboracle.so).  It cannot
// prove the kernels right (warp scans, look-back and shared-memory staging only
// exist on the GPU) but it pins every rule and every boundary case of the
// decomposition -- halos, carries across tiles, ragged windows, word ownership,
// sub-sequence synchronisation -- before GPU minutes are spent.
//
// Build: g++ -O2 -std=c++17 emu.cpp -I<repo> -L<repo>/oracle -loracle
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <random>
#include <string>
#include <vector>

#include "oracle/shafa_oracle.h"
#include "shafa-cd_b200/csrc/core.h"

using namespace shafa;
typedef std::vector<uint8_t> Bytes;

static int g_fail = 0;
#define CHECK(cond, ...)                       \
    do {                                       \
        if (!(cond)) {                         \
            printf("FAIL %s:%d: ", __FILE__, __LINE__); \
            printf(__VA_ARGS__);               \
            printf("\n");                      \
            g_fail++;                          \
            if (g_fail > 20) exit(1);          \
        }                                      \
    } while (0)

// ------------------------------------------------------------------ generators
static Bytes gen(std::mt19937 &rng, size_t n, int kind)
{
    Bytes d(n);
    switch (kind) {
    case 0:  // uniform
        for (auto &b : d) b = rng() & 255;
        break;
    case 1: {  // skewed text-like with zeros
        for (auto &b : d) {
            int r = rng() % 100;
            b = r < 40 ? 'e' : r < 60 ? ' ' : r < 70 ? 0 : (uint8_t)('a' + rng() % 20);
        }
        break;
    }
    case 2: {  // runs, lengths up to 1200, with zeros
        size_t i = 0;
        while (i < n) {
            size_t L = 1 + (rng() % 8 == 0 ? rng() % 1200 : rng() % 6);
            uint8_t s = rng() % 5 == 0 ? 0 : (uint8_t)(rng() % 7);
            for (size_t k = 0; k < L && i < n; k++) d[i++] = s;
        }
        break;
    }
    case 3:  // one giant run with a few interruptions
        std::fill(d.begin(), d.end(), (uint8_t)(rng() % 3 == 0 ? 0 : 9));
        for (int k = 0; k < 5 && n; k++) d[rng() % n] = 1;
        break;
    default:  // two symbols
        for (auto &b : d) b = (rng() & 1) ? 0 : 7;
    }
    return d;
}

// ------------------------------------------------------------------ RLE encode emulation (rle_encode.cu)
// 16 KiB tiles, 64 contiguous bytes per thread, 64-bit masks
static uint32_t rng_noise = 0xa5a5a5a5u;
static Bytes emu_rle_encode_block(const uint8_t *blk, uint64_t len)
{
    Bytes out;
    const int PT = RLE_ITEMS * RLE_VECS;  // 64
    const uint64_t ntiles = (len + RLE_TILE - 1) / RLE_TILE;
    std::vector<uint64_t> run_open(ntiles);  // chain #1: open run length at each tile's end
    for (uint64_t ti = 0; ti < ntiles; ti++) {
        const uint64_t tile_lo = ti * RLE_TILE;
        const uint32_t tile_len = (uint32_t)std::min<uint64_t>(RLE_TILE, len - tile_lo);
        const uint32_t valid_len = (uint32_t)std::min<uint64_t>(RLE_TILE + RLE_HALO, len - tile_lo);
        std::vector<uint32_t> headsv(RLE_HEAD_WORDS + 2, 0);
        uint32_t *heads = headsv.data();
        for (uint32_t p = 0; p < RLE_TILE + RLE_HALO; p++) {
            bool head = (p >= valid_len) || (p == 0 && ti == 0) || (blk[tile_lo + p] != blk[tile_lo + p - 1]);
            if (head) heads[p >> 5] |= 1u << (p & 31);
        }
        int lastv = -1;  // last head among valid tile positions
        for (int p = (int)tile_len - 1; p >= 0; p--)
            if ((heads[p >> 5] >> (p & 31)) & 1) {
                lastv = p;
                break;
            }
        const bool head0 = heads[0] & 1u;
        uint64_t run_in = 0;
        if (lastv >= 0) {
            run_open[ti] = tile_len - lastv;
            if (!head0) run_in = run_open[ti - 1];
        } else {
            run_in = run_open[ti - 1];
            run_open[ti] = run_in + tile_len;
        }
        Bytes stage(2 * RLE_TILE + 64);
        uint32_t o = 0;
        int lh = -1;  // last head before the thread's first position
        auto Ebit = [&](int64_t a) -> uint64_t { return (a > 0 && a < (int64_t)len && blk[a] == blk[a - 1]) ? 1ull : 0ull; };
        for (uint32_t tid = 0; tid < (uint32_t)RLE_THREADS; tid++) {
            const uint32_t p0 = tid * PT;
            const int nvalid = p0 >= tile_len ? 0 : (int)std::min<uint32_t>(PT, tile_len - p0);
            uint64_t lit = 0, tri = 0;
            if (nvalid) {
                const uint64_t valid64 = nvalid >= 64 ? ~0ull : ((1ull << nvalid) - 1ull);
                uint64_t E = 0, Z = 0;
                for (int j = 0; j < nvalid; j++) {
                    E |= Ebit((int64_t)tile_lo + p0 + j) << j;
                    Z |= (uint64_t)(blk[tile_lo + p0 + j] == 0) << j;
                }
                const uint32_t before = (uint32_t)(Ebit((int64_t)tile_lo + p0 - 2) | (Ebit((int64_t)tile_lo + p0 - 1) << 1));
                uint32_t after = 0;
                for (int i = 0; i < 3; i++) after |= (uint32_t)Ebit((int64_t)tile_lo + p0 + 64 + i) << i;
                rng_noise = rng_noise * 1664525u + 1013904223u;
                const uint64_t k0 = lh >= 0 ? (uint64_t)(p0 - lh) : (uint64_t)p0 + run_in;
                const uint64_t my_heads = ~E;
                // the general rule, group by group (this is also the kernel's slow path)
                uint64_t glit = 0, gtri = 0;
                for (int c = 0; c < RLE_VECS; c++) {
                    const int nv = nvalid - 16 * c;
                    if (nv <= 0) break;
                    const uint64_t hb = c == 0 ? 0ull : (my_heads & ((1ull << (16 * c)) - 1ull));
                    uint64_t kc;
                    if (hb)
                        kc = 16 * c - (63 - __builtin_clzll(hb));
                    else
                        kc = ((E & 1ull) ? k0 : 0) + 16 * c;  // if p0 is a head, hb is non-zero for c > 0 and k = 0 for c = 0
                    auto symc = [&](int j) -> uint32_t { return blk[tile_lo + p0 + 16 * c + j]; };
                    const RleEmit em = rle_emit_thread(heads, p0 + 16 * c, std::min(nv, 16), (uint32_t)(kc % 255), symc);
                    glit |= (uint64_t)em.lit << (16 * c);
                    gtri |= (uint64_t)em.tri << (16 * c);
                }
                {
                    uint64_t Ev = E;
                    uint32_t bv = before, av = after;
                    if ((E & 1ull) && k0 + (PT + 2) >= 255) rle_chunk_heads(k0, &Ev, &bv, &av);  // as in the kernel
                    const RleMasks64 m = rle_fast_masks64(Ev, bv | (rng_noise & ~3u), av | (rng_noise & ~7u), Z & valid64, valid64);
                    lit = m.lit;
                    tri = m.tri;
                    CHECK(lit == glit && tri == gtri, "fast masks differ from the general rule (tile %llu tid %u k0 %llu)",
                          (unsigned long long)ti, tid, (unsigned long long)k0);
                }
            }
            for (int j = 0; j < PT; j++) {
                const uint32_t sy = (j < nvalid) ? blk[tile_lo + p0 + j] : 0;
                if ((tri >> j) & 1) {
                    stage[o++] = 0;
                    stage[o++] = (uint8_t)sy;
                    stage[o++] = (uint8_t)rle_run_remaining(heads, p0 + j);
                } else if ((lit >> j) & 1) {
                    stage[o++] = (uint8_t)sy;
                }
            }
            for (uint32_t p = p0; p < p0 + (uint32_t)PT; p++)
                if ((heads[p >> 5] >> (p & 31)) & 1) lh = (int)p;
        }
        out.insert(out.end(), stage.begin(), stage.begin() + o);
    }
    return out;
}

static void test_rle_encode(std::mt19937 &rng)
{
    const size_t sizes[] = {1, 2, 15, 16, 17, 63, 64, 65, 254, 255, 256, 4095, 4096, 4097, 16383, 16384, 16385, 16384 + 255, 32768 + 3, 40000, 70000};
    for (int kind = 0; kind < 5; kind++)
        for (size_t n : sizes) {
            Bytes d = gen(rng, n + 1, kind);
            d.insert(d.begin(), 0x55);  // guard byte before the block (blk[-1] is read for ti>0 only)
            const uint8_t *blk = d.data() + 1;
            Bytes want(2 * n + 3);
            want.resize(orc_rle_encode_block(blk, n, want.data()));
            Bytes got = emu_rle_encode_block(blk, n);
            CHECK(got == want, "rle_encode kind %d n %zu: got %zu want %zu bytes", kind, n, got.size(), want.size());
        }
}

// ------------------------------------------------------------------ RLE decode emulation (rle_decode.cu)
// stream = concatenated ragged blocks; windows are 16 KiB-aligned in stream space, 64 bytes per thread
static Bytes emu_rle_decode(const Bytes &stream, const std::vector<uint64_t> &off, std::vector<uint64_t> &fin, bool &trunc)
{
    Bytes out;
    trunc = false;
    const int T = 16384, PT = 64, NT = 256;
    for (size_t b = 0; b + 1 < off.size(); b++) {
        const uint64_t lo = off[b], hi = off[b + 1];
        const uint64_t out_before = out.size();
        if (hi == lo) {
            fin.push_back(0);
            continue;
        }
        uint32_t tile_entry = 0;
        for (uint64_t w = lo / T; w <= (hi - 1) / T; w++) {
            const uint64_t base = w * T;
            std::vector<uint32_t> maps(NT);
            std::vector<int> J0(NT), N(NT);
            std::vector<uint64_t> ZZ(NT);
            for (int tid = 0; tid < NT; tid++) {
                const uint64_t pos = base + (uint64_t)tid * PT;
                int j0 = 0, j1 = 0;
                if (pos + PT > lo && pos < hi) {
                    j0 = lo > pos ? (int)(lo - pos) : 0;
                    j1 = hi < pos + PT ? (int)(hi - pos) : PT;
                }
                uint64_t z = 0;
                for (int j = j0; j < j1; j++)
                    if (stream[pos + j] == 0) z |= 1ull << (j - j0);
                J0[tid] = j0;
                N[tid] = j1 - j0;
                ZZ[tid] = z;
                maps[tid] = fsm_map64(z, j1 - j0);
            }
            uint32_t excl = FSM_IDENTITY, st = tile_entry;
            for (int tid = 0; tid < NT; tid++) {
                const uint64_t pos = base + (uint64_t)tid * PT;
                const uint32_t my_entry = fsm_apply(excl, tile_entry);
                CHECK(my_entry == st, "fsm map scan disagrees with sequential state (tid %d)", tid);
                const int n = N[tid], j0 = J0[tid];
                const FsmWalk wk = fsm_walk64(ZZ[tid], n, my_entry);
                const RleDecMasks m = rle_dec_masks64(wk.esc, n, my_entry);
                // sequential reference for this thread's bytes
                {
                    uint32_t s = my_entry;
                    uint64_t rlit = 0, rrun = 0;
                    for (int q = 0; q < n; q++) {
                        const uint8_t v = stream[pos + j0 + q];
                        if (s == 0) {
                            if (v == 0) s = 1; else rlit |= 1ull << q;
                        } else if (s == 1) s = 2;
                        else { rrun |= 1ull << q; s = 0; }
                    }
                    CHECK(rlit == m.lit && rrun == m.run && s == wk.exit, "decode masks/exit differ from the sequential automaton (tid %d n %d entry %u)", tid, n, my_entry);
                    st = s;
                }
                for (int q = 0; q < n; q++) {
                    const uint8_t cur = stream[pos + j0 + q];
                    if ((m.run >> q) & 1) {
                        const uint8_t sym = stream[pos + j0 + q - 1];
                        out.insert(out.end(), cur ? cur : 1, sym);
                    } else if ((m.lit >> q) & 1)
                        out.push_back(cur);
                }
                excl = fsm_compose(excl, maps[tid]);
            }
            CHECK(fsm_apply(excl, tile_entry) == st, "tile map exit mismatch");
            tile_entry = st;
        }
        if (tile_entry != 0) trunc = true;
        fin.push_back(out.size() - out_before);
    }
    return out;
}

static void test_rle_decode(std::mt19937 &rng)
{
    for (int rep = 0; rep < 40; rep++) {
        // several ragged blocks of valid RLE (made by the oracle encoder) at arbitrary offsets
        Bytes stream, want;
        std::vector<uint64_t> off{0};
        const int nblk = 1 + rng() % 4;
        if (rep % 3 == 0) stream.resize(rng() % 5000, 0xEE), off[0] = stream.size();  // unaligned start
        for (int b = 0; b < nblk; b++) {
            Bytes raw = gen(rng, 1 + rng() % 30000, rng() % 5);
            Bytes r(2 * raw.size() + 3);
            r.resize(orc_rle_encode_block(raw.data(), raw.size(), r.data()));
            stream.insert(stream.end(), r.begin(), r.end());
            want.insert(want.end(), raw.begin(), raw.end());
            off.push_back(stream.size());
        }
        stream.resize(stream.size() + 64, 0xEE);
        std::vector<uint64_t> fin;
        bool trunc;
        Bytes got = emu_rle_decode(stream, off, fin, trunc);
        CHECK(!trunc, "rle_decode reported truncation on valid data");
        CHECK(got == want, "rle_decode rep %d: got %zu want %zu", rep, got.size(), want.size());
    }
    // arbitrary (not encoder-made) streams incl. count 0 and zero syms: compare with the oracle decoder
    for (int rep = 0; rep < 60; rep++) {
        Bytes s(1 + rng() % 9000);
        for (auto &b : s) b = (rng() % 3 == 0) ? 0 : (uint8_t)(rng() % 4);
        uint64_t fin_o = 0;
        int rc = orc_rle_decode_block(s.data(), s.size(), nullptr, 0, &fin_o);
        Bytes want(fin_o);
        if (rc == 0) rc = orc_rle_decode_block(s.data(), s.size(), want.data(), want.size(), &fin_o);
        Bytes st = s;
        st.resize(st.size() + 64);
        std::vector<uint64_t> fin;
        bool trunc;
        Bytes got = emu_rle_decode(st, {0, s.size()}, fin, trunc);
        if (rc == 0) {
            CHECK(!trunc && got == want, "rle_decode arbitrary rep %d", rep);
        } else {
            CHECK(trunc, "oracle rejects but emulation did not flag truncation (rep %d)", rep);
        }
    }
}

// ------------------------------------------------------------------ SF build emulation (sf_build.cu)
static int emu_sf_build(const uint64_t *hist, uint8_t *len, uint64_t *code)
{
    uint32_t cnt[256], sorted[256], symr[256], P[257];
    for (int s = 0; s < 256; s++) cnt[s] = (uint32_t)hist[s];
    for (int s = 0; s < 256; s++) {
        uint32_t rank = 0;
        for (int t = 0; t < 256; t++) rank += (cnt[t] > cnt[s]) || (cnt[t] == cnt[s] && t < s);
        sorted[rank] = cnt[s];
        symr[rank] = s;
    }
    P[0] = 0;
    int nz = 0;
    for (int r = 0; r < 256; r++) {
        P[r + 1] = P[r] + sorted[r];
        nz += sorted[r] != 0;
    }
    const int nn = nz - 1;
    int lo[256], hi[256];
    uint64_t cd[256] = {0};
    uint32_t ln[256] = {0};
    for (int r = 0; r < 256; r++) lo[r] = 0, hi[r] = nn;
    int maxlen = 0;
    for (;;) {
        bool any = false;
        int d[256];
        for (int r = 0; r <= nn; r++)
            if (lo[r] != hi[r]) {
                any = true;
                d[r] = sf_split(P, lo[r], hi[r]);
            }
        if (!any) break;
        for (int r = 0; r <= nn; r++)
            if (lo[r] != hi[r]) {
                if (r <= d[r]) {
                    hi[r] = d[r];
                    cd[r] <<= 1;
                } else {
                    lo[r] = d[r] + 1;
                    cd[r] = (cd[r] << 1) | 1;
                }
                ln[r]++;
            }
    }
    memset(len, 0, 256);
    memset(code, 0, 256 * 8);
    for (int r = 0; r <= nn; r++) {
        len[symr[r]] = (uint8_t)ln[r];
        code[symr[r]] = cd[r];
        maxlen = std::max(maxlen, (int)ln[r]);
    }
    return maxlen;
}

static void test_sf_build(std::mt19937 &rng)
{
    for (int rep = 0; rep < 3000; rep++) {
        uint64_t h[256] = {0};
        const int kind = rep % 6;
        const int k = 1 + rng() % 256;
        for (int i = 0; i < k; i++) {
            const int s = rng() % 256;
            switch (kind) {
            case 0: h[s] = 1 + rng() % 1000000; break;
            case 1: h[s] = 1 + rng() % 3; break;
            case 2: h[s] = 5; break;
            case 3: h[s] = 1ull << (rng() % 20); break;
            case 4: h[s] = 1 + rng() % 40; break;
            default: h[s] = 2500 + rng() % 100;
            }
        }
        if (rep == 7) {  // fibonacci chain
            memset(h, 0, sizeof h);
            uint64_t a = 1, b = 1;
            for (int i = 0; i < 38; i++) {
                h[100 + i] = a;
                uint64_t c = a + b;
                a = b;
                b = c;
            }
        }
        uint8_t l1[256], l2[256];
        uint64_t c1[256], c2[256];
        const int m1 = orc_sf_codes_packed(h, l1, c1);
        const int m2 = emu_sf_build(h, l2, c2);
        CHECK(m1 == m2 && !memcmp(l1, l2, 256) && !memcmp(c1, c2, sizeof c1), "sf_build rep %d (max %d vs %d)", rep, m1, m2);
    }
}

// ------------------------------------------------------------------ SF encode emulation (sf_encode.cu fast path)
// 16 KiB tiles, 64 contiguous symbols per thread, bit range materialised window by window
static Bytes emu_sf_encode(const Bytes &stream, uint64_t lo, uint64_t hi, const uint8_t *len, const uint64_t *code,
                           uint32_t stage_words = ENC_STAGE_WORDS)
{
    const int T = ENC_TILE, NT = ENC_THREADS, IT = ENC_ITEMS;
    uint32_t lut[256];
    for (int s = 0; s < 256; s++) lut[s] = enc_entry(code[s], len[s]);
    uint64_t total_bits = 0;
    for (uint64_t p = lo; p < hi; p++) total_bits += len[stream[p]];
    std::vector<uint32_t> outw((total_bits + 31) / 32 + 4, 0xdeadbeef);  // detect unwritten / doubly written words
    std::vector<int> written(outw.size(), 0);
    uint64_t P = 0;
    const uint64_t w_first = lo / T, w_last = (hi - 1) / T;
    for (uint64_t w = w_first; w <= w_last; w++) {
        const uint64_t base = w * T;
        const bool is_first = w == w_first, is_last = w == w_last;
        std::vector<uint32_t> bits(NT, 0);
        uint32_t tile_bits = 0;
        for (int tid = 0; tid < NT; tid++) {
            for (int k = 0; k < IT; k++) {
                const uint64_t q = base + (uint64_t)tid * IT + k;
                if (q >= lo && q < hi) bits[tid] += enc_len(lut[stream[q]]);
            }
            tile_bits += bits[tid];
        }
        const uint64_t E = P + tile_bits, W0 = P >> 5, Wtouch = (E + 31) >> 5, Wown = is_last ? Wtouch : (E >> 5);
        for (uint64_t win = W0; win < Wtouch; win += stage_words) {
            const uint32_t nw = (uint32_t)std::min<uint64_t>(Wtouch - win, stage_words);
            std::vector<uint32_t> stage(nw, 0);
            uint64_t excl = 0;
            for (int tid = 0; tid < NT; tid++) {
                const uint64_t p = P + excl;
                const uint64_t wf = p >> 5, wl = (p + bits[tid] + 31) >> 5;
                if (bits[tid] && wl > win && wf < win + nw) {
                    int32_t o = (int32_t)((int64_t)p - (int64_t)(win << 5));
                    for (int k = 0; k < IT; k++) {
                        const uint64_t q64 = base + (uint64_t)tid * IT + k;
                        const uint32_t ee = (q64 >= lo && q64 < hi) ? lut[stream[q64]] : 0;
                        const uint32_t l = enc_len(ee), q = (uint32_t)o & 31u, wi = (uint32_t)(o >> 5);
                        const uint64_t v64 = (uint64_t)enc_code(ee) << ((64u - q - l) & 63u);
                        const uint32_t hiw = (uint32_t)(v64 >> 32), low = (uint32_t)v64;
                        if (hiw && wi < nw) {
                            CHECK((stage[wi] & hiw) == 0, "overlapping bits in a staging word");
                            stage[wi] |= hiw;
                        }
                        if (low && wi + 1u < nw) {
                            CHECK((stage[wi + 1] & low) == 0, "overlapping bits in a staging word");
                            stage[wi + 1] |= low;
                        }
                        o += (int32_t)l;
                    }
                }
                excl += bits[tid];
            }
            if (win == W0 && !is_first && (P & 31)) {  // halo: lanes 0..31 look at the symbols before the tile
                const uint32_t r = P & 31;
                uint32_t s_prev = 0, wv = 0;
                for (int lane = 0; lane < 32; lane++) {
                    uint32_t ee = 0;
                    if (base >= lo + 1 + lane) ee = lut[stream[base - 1 - lane]];
                    const uint32_t l = enc_len(ee);
                    if (l && s_prev < r) {
                        const uint32_t room = r - s_prev;
                        wv |= (enc_code(ee) & ((1u << room) - 1u)) << (32 - r + s_prev);
                    }
                    s_prev += l;
                }
                stage[0] |= wv;
            }
            const uint64_t cend = std::min<uint64_t>(Wown, win + nw);
            for (uint64_t ww = win; ww < cend; ww++) {
                CHECK(!written[ww], "output word %llu written twice", (unsigned long long)ww);
                written[ww] = 1;
                outw[ww] = stage[ww - win];
            }
        }
        P = E;
    }
    Bytes out((total_bits + 7) / 8);
    for (size_t i = 0; i < out.size(); i++) {
        CHECK(written[i / 4], "output word %zu never written", i / 4);
        out[i] = (uint8_t)(outw[i / 4] >> (24 - 8 * (i & 3)));
    }
    return out;
}

static void codes_to_strings(const uint8_t *len, const uint64_t *code, char (*out)[ORC_MAX_CODE])
{
    memset(out, 0, 256 * ORC_MAX_CODE);
    for (int s = 0; s < 256; s++)
        for (int i = 0; i < len[s]; i++) out[s][i] = ((code[s] >> (len[s] - 1 - i)) & 1) ? '1' : '0';
}

// ------------------------------------------------------------------ SF decode emulation (sf_decode.cu)
struct EmuTables {
    std::vector<uint16_t> lut;
    std::vector<uint8_t> lut2;
    std::vector<uint32_t> lut3;
    std::vector<uint8_t> lut16;  // only for "lean" codes (complete, <= 12 bits): all whole codes of a 16-bit window
    bool lean = false;
    uint64_t sleft[256];
    uint8_t slen[256], ssym[256];
    DecView view(const uint64_t *pairs) const
    {
        DecView t;
        t.pairs = pairs;
        t.lut1 = lut.data();
        t.lut2 = lut2.data();
        t.lut3 = lut3.data();
        t.lut16 = lut16.data();
        t.sleft = sleft;
        t.slen = slen;
        t.ssym = ssym;
        return t;
    }
};

static void emu_dec_tables(const uint8_t *len, const uint64_t *code, EmuTables &T)
{
    uint64_t left[256];
    for (int s = 0; s < 256; s++) left[s] = len[s] ? code[s] << (64 - len[s]) : ~0ull;
    for (int s = 0; s < 256; s++) {
        uint32_t rank = 0;
        for (int t = 0; t < 256; t++) {
            const bool before = (left[t] < left[s]) ||
                                (left[t] == left[s] && ((len[t] != 0 && len[s] == 0) || ((len[t] != 0) == (len[s] != 0) && t < s)));
            rank += before;
        }
        T.sleft[rank] = left[s];
        T.slen[rank] = len[s];
        T.ssym[rank] = (uint8_t)s;
    }
    T.lut.assign(DEC_LUT_SIZE, 0);
    for (int e = 0; e < DEC_LUT_SIZE; e++) {
        const uint64_t window = (uint64_t)e << (64 - DEC_LUT_BITS);
        int lo = 0, hi = 256;
        while (lo < hi) {
            const int m = (lo + hi) >> 1;
            if (T.sleft[m] <= window)
                lo = m + 1;
            else
                hi = m;
        }
        uint16_t entry = 0;
        if (lo > 0) {
            const int r = lo - 1;
            const uint32_t rl = T.slen[r];
            if (rl != 0 && rl <= (uint32_t)DEC_LUT_BITS && (window >> (64 - rl)) == (T.sleft[r] >> (64 - rl)))
                entry = (uint16_t)((T.ssym[r] << 8) | rl);
            else
                entry = DEC_LONG;
        }
        T.lut[e] = entry;
    }
    T.lut2.assign(DEC_LUT2_SIZE, 0);
    T.lut3.assign(DEC_LUT_SIZE, 0);
    for (int e = 0; e < DEC_LUT_SIZE; e++) {
        T.lut2[2 * e] = dec_lut2_entry(T.lut.data(), 2u * (uint32_t)e);
        T.lut2[2 * e + 1] = dec_lut2_entry(T.lut.data(), 2u * (uint32_t)e + 1u);
        T.lut3[e] = dec_lut3_entry(T.lut.data(), (uint32_t)e);
    }
    // lean: every 12-bit window holds a whole code (what k_dec_sync checks before it trusts lut1 blindly)
    T.lean = true;
    for (int e = 0; e < DEC_LUT_SIZE; e++)
        if ((uint32_t)(T.lut[e] & 0xFF) - 1u >= (uint32_t)DEC_LUT_BITS) T.lean = false;
    T.lut16.assign(DEC_LUT16_SIZE, 0);
    if (T.lean)
        for (int v = 0; v < DEC_LUT16_SIZE; v++) T.lut16[v] = dec_lut16_entry(T.lut.data(), (uint32_t)v);
}

static long g_lean_walks = 0;
static Bytes emu_sf_decode(const Bytes &payload, uint64_t n, const uint8_t *len, const uint64_t *code, bool &bad, int *stitch_passes)
{
    EmuTables T;
    emu_dec_tables(len, code, T);
    const uint64_t z = payload.size();
    const uint64_t nseq = (z + DEC_SEQ_BYTES - 1) / DEC_SEQ_BYTES;
    std::vector<uint16_t> start(nseq * DEC_THREADS), count(nseq * DEC_THREADS);
    std::vector<uint32_t> seq_end(nseq);
    std::vector<uint32_t> wordsv(DEC_THREADS * DEC_ROW_WORDS);  // the same sub-sequences as register words (dec_*_words)
    auto load_pairs = [&](uint64_t seq, std::vector<uint64_t> &pairs) {  // dec_load_rows: one row per sub-sequence
        pairs.assign(DEC_PAIR_SLOTS, 0);
        for (int sid = 0; sid < DEC_THREADS; sid++) {
            uint32_t words[DEC_ROW_WORDS];
            for (int i = 0; i < DEC_ROW_WORDS; i++) {
                uint32_t w = 0;
                for (int k = 0; k < 4; k++) {
                    const uint64_t bo = seq * DEC_SEQ_BYTES + (uint64_t)sid * 16 + (uint64_t)i * 4 + k;
                    w = (w << 8) | (bo < z ? payload[bo] : 0);
                }
                words[i] = w;
                wordsv[(size_t)sid * DEC_ROW_WORDS + i] = w;
            }
            dec_row_from_words(words, pairs.data() + (size_t)sid * DEC_ROW);
        }
    };
    auto words_of = [&](int sid) -> const uint32_t (&)[DEC_ROW_WORDS] {
        return *reinterpret_cast<const uint32_t (*)[DEC_ROW_WORDS]>(wordsv.data() + (size_t)sid * DEC_ROW_WORDS);
    };
    auto sub_stop = [](uint32_t limit, uint32_t sub_lo) -> uint32_t {
        return limit > sub_lo ? std::min<uint32_t>(limit - sub_lo, DEC_SUB_BITS) : 0u;
    };
    // phase 1
    for (uint64_t seq = 0; seq < nseq; seq++) {
        std::vector<uint64_t> pairs;
        load_pairs(seq, pairs);
        const DecView t = T.view(pairs.data());
        const uint64_t bits_left = (z - seq * DEC_SEQ_BYTES) * 8;
        const uint32_t limit = (uint32_t)std::min<uint64_t>(bits_left, DEC_SEQ_BYTES * 8);
        uint32_t st[DEC_THREADS] = {0}, en[DEC_THREADS], cn[DEC_THREADS];
        auto run = [&](int tid) {
            const uint32_t stop = sub_stop(limit, tid * DEC_SUB_BITS);
            DecSub r = dec_count_subsequence(t, tid * DEC_ROW, st[tid], stop);
            DecSub r1 = dec_emit_subsequence(t, tid * DEC_ROW, st[tid], stop, DecNoEmit());
            CHECK(r.end == r1.end && r.count == r1.count, "multi-symbol and single-symbol walks disagree");
            const DecSub rw = dec_count_words(t, words_of(tid), st[tid], stop);
            CHECK(rw.end == r.end && rw.count == r.count && rw.bad == r.bad, "register-window count walk disagrees with the row walk");
            en[tid] = r.end > (uint32_t)DEC_SUB_BITS ? r.end - DEC_SUB_BITS : 0;
            cn[tid] = r.count;
            if (T.lean) {  // the multi-code walk of k_dec_sync_lean, from this start and from every other one it would try
                for (uint32_t s2 = 0; s2 < 15; s2 += 3) {
                    const uint32_t st2 = s2 == 12 ? st[tid] : s2;
                    uint32_t c3[3], e3[3];
                    dec_lean_walk<3>(t, words_of(tid), st2, stop, c3, e3);
                    for (uint32_t k = 0; k < 3; k++) {
                        const DecSub r2 = dec_count_subsequence(t, tid * DEC_ROW, st2 + k, stop);
                        CHECK(c3[k] == r2.count && e3[k] == r2.end, "lean multi-code walk differs from the plain walk");
                        g_lean_walks++;
                    }
                }
            }
        };
        for (int tid = 0; tid < DEC_THREADS; tid++) run(tid);
        for (int iter = 0;; iter++) {
            uint32_t want[DEC_THREADS];
            bool any = false;
            for (int tid = 0; tid < DEC_THREADS; tid++) want[tid] = tid ? en[tid - 1] : 0;
            for (int tid = 0; tid < DEC_THREADS; tid++)
                if (want[tid] != st[tid]) {
                    any = true;
                    st[tid] = want[tid];
                    run(tid);
                }
            if (!any) break;
            CHECK(iter < 300, "phase 1 does not converge");
        }
        for (int tid = 0; tid < DEC_THREADS; tid++) {
            CHECK(st[tid] < 128 && cn[tid] < 512, "sub record overflow");
            start[seq * DEC_THREADS + tid] = (uint16_t)st[tid];
            count[seq * DEC_THREADS + tid] = (uint16_t)cn[tid];
        }
        seq_end[seq] = en[DEC_THREADS - 1];
    }
    // phase 2
    int passes = 0;
    for (;; passes++) {
        bool touched = false;
        for (uint64_t seq = 1; seq < nseq; seq++) {
            uint32_t cur = seq_end[seq - 1];
            std::vector<uint64_t> pairs;
            load_pairs(seq, pairs);
            const DecView t = T.view(pairs.data());
            const uint64_t bits_left = (z - seq * DEC_SEQ_BYTES) * 8;
            const uint32_t limit = (uint32_t)std::min<uint64_t>(bits_left, DEC_SEQ_BYTES * 8);
            for (int k = 0; k < DEC_THREADS; k++) {
                const uint64_t gi = seq * DEC_THREADS + k;
                if (start[gi] == cur) break;
                DecSub r = dec_count_subsequence(t, k * DEC_ROW, cur, sub_stop(limit, k * DEC_SUB_BITS));
                start[gi] = (uint16_t)cur;
                count[gi] = (uint16_t)r.count;
                touched = true;
                cur = r.end > (uint32_t)DEC_SUB_BITS ? r.end - DEC_SUB_BITS : 0;
                if (k == DEC_THREADS - 1) seq_end[seq] = cur;
            }
        }
        if (!touched) break;
        CHECK(passes < 1000, "phase 2 does not converge");
    }
    if (stitch_passes) *stitch_passes = passes;
    // phases 3+4
    Bytes out(n);
    uint64_t first = 0;
    bad = false;
    for (uint64_t seq = 0; seq < nseq; seq++) {
        std::vector<uint64_t> pairs;
        load_pairs(seq, pairs);
        const DecView t = T.view(pairs.data());
        const uint64_t bits_left = (z - seq * DEC_SEQ_BYTES) * 8;
        const uint32_t limit = (uint32_t)std::min<uint64_t>(bits_left, DEC_SEQ_BYTES * 8);
        for (int tid = 0; tid < DEC_THREADS; tid++) {
            const uint64_t gi = seq * DEC_THREADS + tid;
            if (first < n) {
                const uint64_t room = n - first;
                uint8_t *dst = out.data() + first;
                DecSub r = dec_emit_subsequence(t, tid * DEC_ROW, start[gi], sub_stop(limit, tid * DEC_SUB_BITS),
                                                [&](uint32_t i, uint32_t sym) { if (i < room) dst[i] = (uint8_t)sym; });
                CHECK(r.count == count[gi], "final decode count differs from the recorded count");
                {
                    Bytes alt(r.count + 2);
                    auto put = [&](uint32_t i, uint32_t sym) { alt[i] = (uint8_t)sym; };
                    DecFnSink<decltype(put)> sink{put, 0u};
                    const DecSub rw = dec_emit_words(t, words_of(tid), start[gi], sub_stop(limit, tid * DEC_SUB_BITS), sink);
                    Bytes ref(r.count + 2);
                    dec_emit_subsequence(t, tid * DEC_ROW, start[gi], sub_stop(limit, tid * DEC_SUB_BITS), [&](uint32_t i, uint32_t sym) { ref[i] = (uint8_t)sym; });
                    CHECK(rw.end == r.end && rw.count == r.count && alt == ref, "register-window emit walk disagrees with the row walk");
                }
                if (r.bad) bad = true;
            }
            first += count[gi];
        }
    }
    if (first < n) bad = true;
    return out;
}


// ------------------------------------------------------------------ fused single-pass decoder emulation (sf_decode_fused.cu)
// Same control flow as k_dec_fused, sequence after sequence: speculative pass from bit 0 into record lists, rounds with
// the assumed carry 0, "publish" the end, rounds with the true carry, verify the published end, compaction.
struct EmuFusedView {
    const uint64_t *pairs;
    const uint32_t *lutp;
    const uint64_t *sleft;
    const uint8_t *slen, *ssym;
    uint32_t *slots;   // [sub-sequence][DECF_CAP], addressed in bytes
    uint8_t *heads;    // [sub-sequence][DECF_HEAD]
    uint64_t pair(uint32_t slot) const { return pairs[slot]; }
    uint32_t lp(uint32_t i) const { return lutp[i]; }
    uint32_t ld32(uint32_t a) const { return slots[a / 4]; }
    void st32(uint32_t a, uint32_t v) const { slots[a / 4] = v; }
    void st8(uint32_t a, uint32_t v) const { heads[a] = (uint8_t)v; }
};
struct EmuFusedSub {
    DecfMain mn;
    uint32_t start, end, count, hn, jsp;
};
static int g_fused_specfail = 0, g_fused_kind[3] = {0, 0, 0};
static Bytes emu_sf_decode_fused(const Bytes &payload, uint64_t n, const uint8_t *len, const uint64_t *code, bool &bad)
{
    EmuTables T;
    emu_dec_tables(len, code, T);
    std::vector<uint32_t> lutp(DEC_LUT_SIZE);
    for (int e = 0; e < DEC_LUT_SIZE; e++) lutp[e] = decf_entry(T.lut.data(), (uint32_t)e);
    const uint64_t z = payload.size();
    const uint64_t nseq = (z + DEC_SEQ_BYTES - 1) / DEC_SEQ_BYTES;
    Bytes out(n);
    uint64_t first = 0;
    uint32_t prev_end = 0;
    bad = false;
    std::vector<uint32_t> slots(DEC_THREADS * DECF_CAP);
    std::vector<uint8_t> heads(DEC_THREADS * DECF_HEAD);
    for (uint64_t seq = 0; seq < nseq; seq++) {
        std::vector<uint64_t> pairs(DEC_PAIR_SLOTS, 0);
        for (int sid = 0; sid < DEC_THREADS; sid++) {
            uint32_t words[DEC_ROW_WORDS];
            for (int i = 0; i < DEC_ROW_WORDS; i++) {
                uint32_t w = 0;
                for (int k = 0; k < 4; k++) {
                    const uint64_t bo = seq * DEC_SEQ_BYTES + (uint64_t)sid * 16 + (uint64_t)i * 4 + k;
                    w = (w << 8) | (bo < z ? payload[bo] : 0);
                }
                words[i] = w;
            }
            dec_row_from_words(words, pairs.data() + (size_t)sid * DEC_ROW);
        }
        EmuFusedView t{pairs.data(), lutp.data(), T.sleft, T.slen, T.ssym, slots.data(), heads.data()};
        const uint64_t bits_left = (z - seq * DEC_SEQ_BYTES) * 8;
        const uint32_t limit = (uint32_t)std::min<uint64_t>(bits_left, DEC_SEQ_BYTES * 8);
        auto stop_of = [&](int tid) -> uint32_t {
            const uint32_t sub_lo = tid * DEC_SUB_BITS;
            return limit > sub_lo ? std::min<uint32_t>(limit - sub_lo, DEC_SUB_BITS) : 0u;
        };
        auto over = [](uint32_t end) { return end > (uint32_t)DEC_SUB_BITS ? end - DEC_SUB_BITS : 0u; };
        EmuFusedSub u[DEC_THREADS];
        uint32_t ends[DEC_THREADS];
        for (int tid = 0; tid < DEC_THREADS; tid++) {
            u[tid].mn = decf_pass(t, tid * DEC_ROW, tid * DECF_CAP * 4, 0u, stop_of(tid));
            u[tid].start = 0;
            u[tid].hn = u[tid].jsp = 0;
            u[tid].count = u[tid].mn.k;
            u[tid].end = u[tid].mn.end;
            ends[tid] = over(u[tid].end);
            if (u[tid].mn.bad) bad = true;
        }
        auto restart = [&](int tid, uint32_t s) {
            EmuFusedSub &x = u[tid];
            const uint32_t row = tid * DEC_ROW, slot = tid * DECF_CAP * 4, head = tid * DECF_HEAD, stop = stop_of(tid);
            x.start = s;
            const DecfSplice sp = decf_splice(t, row, slot, head, x.mn, s, stop);
            g_fused_kind[sp.kind]++;
            const uint32_t mrec = x.mn.norec ? 0u : x.mn.m;
            if (sp.kind == 0) {
                x.hn = sp.hn; x.jsp = sp.jsp; x.count = sp.hn + (x.mn.k - sp.kskip); x.end = x.mn.end;
            } else if (sp.kind == 1) {
                x.hn = sp.hn; x.jsp = mrec; x.count = sp.hn; x.end = sp.end;
            } else {
                x.mn = decf_pass(t, row, slot, s, stop);
                x.hn = 0; x.jsp = 0; x.count = x.mn.k; x.end = x.mn.end;
            }
        };
        auto rounds = [&](uint32_t carry) {
            for (int iter = 0;; iter++) {
                uint32_t want[DEC_THREADS];
                for (int tid = 0; tid < DEC_THREADS; tid++) want[tid] = tid ? ends[tid - 1] : carry;   // snapshot (barrier)
                bool any = false;
                for (int tid = 0; tid < DEC_THREADS; tid++)
                    if (want[tid] != u[tid].start) {
                        any = true;
                        restart(tid, want[tid]);
                        ends[tid] = over(u[tid].end);
                    }
                if (!any) break;
                CHECK(iter < 600, "fused rounds do not converge");
            }
        };
        rounds(0);
        const uint32_t end0 = ends[DEC_THREADS - 1];
        if (seq != 0) {
            rounds(prev_end);
            if (ends[DEC_THREADS - 1] != end0) g_fused_specfail++;
        }
        prev_end = ends[DEC_THREADS - 1];
        // counts, compaction (stage), copy-out clipped at n
        uint32_t total = 0;
        bool direct = false;
        for (int tid = 0; tid < DEC_THREADS; tid++) {
            total += u[tid].count;
            direct |= u[tid].mn.norec != 0;
        }
        direct |= total > (uint32_t)(DEC_PAIR_SLOTS * 8);
        Bytes stage(total + 2);
        uint32_t d = 0;
        for (int tid = 0; tid < DEC_THREADS; tid++) {
            const uint32_t d0 = d;
            if (!direct) {
                for (uint32_t i = 0; i < u[tid].hn; i++) stage[d++] = heads[tid * DECF_HEAD + i];
                const uint32_t mrec = u[tid].mn.norec ? 0u : u[tid].mn.m;
                for (uint32_t j = u[tid].jsp; j < mrec; j++) {
                    const uint32_t r = slots[tid * DECF_CAP + j];
                    stage[d] = (uint8_t)r;
                    if (r >> 29) stage[d + 1] = (uint8_t)(r >> 8);
                    d += r >> 28;
                }
            } else {
                uint32_t q = u[tid].start, b2 = 0;
                const uint32_t stop = stop_of(tid);
                while (q < stop) {
                    const uint32_t e = decf_lookup(t, tid * DEC_ROW, q, &b2);
                    stage[d++] = (uint8_t)e;
                    q += decf_n(e) == 2u ? decf_len0(e) : decf_total(e);
                }
            }
            CHECK(d - d0 == u[tid].count, "fused: compaction wrote %u symbols, count says %u (seq %llu sub %d)", d - d0, u[tid].count,
                  (unsigned long long)seq, tid);
        }
        for (uint32_t i = 0; i < total && first + i < n; i++) out[first + i] = stage[i];
        first += total;
    }
    if (first < n) bad = true;
    return out;
}

static void test_sf_codec(std::mt19937 &rng)
{
    int max_passes = 0;
    for (int rep = 0; rep < 60; rep++) {
        const int kind = rep % 5;
        size_t n = 1 + rng() % 60000;
        if (rep % 7 == 0) n = 1 + rng() % 40;
        Bytes raw = gen(rng, n, kind);
        if (rep % 11 == 3) {  // deep code: fibonacci-weighted symbols -> long codes (fast path needs <= 26 bits)
            uint64_t a = 1, b = 1;
            raw.clear();
            for (int i = 0; i < 22; i++) {
                raw.insert(raw.end(), (size_t)a, (uint8_t)(10 + i));
                uint64_t c = a + b;
                a = b;
                b = c;
            }
            std::shuffle(raw.begin(), raw.end(), rng);
            n = raw.size();
        }
        uint64_t hist[256];
        orc_histogram(raw.data(), n, hist);
        int distinct = 0;
        for (int s = 0; s < 256; s++) distinct += hist[s] != 0;
        if (distinct < 2) continue;
        uint8_t len[256];
        uint64_t code[256];
        const int maxlen = orc_sf_codes_packed(hist, len, code);
        CHECK(maxlen > 0 && maxlen <= ENC_SHORT_MAX, "unexpected max code length %d", maxlen);
        static char codes[256][ORC_MAX_CODE];
        codes_to_strings(len, code, codes);
        Bytes want((n * maxlen + 7) / 8 + 8);
        want.resize(orc_sf_encode_block(raw.data(), n, codes, want.data()));
        // embed at an unaligned offset inside a larger stream
        const uint64_t lo = rng() % 9000;
        Bytes stream(lo, 0xAA);
        stream.insert(stream.end(), raw.begin(), raw.end());
        stream.resize(stream.size() + 64, 0xBB);
        Bytes got = emu_sf_encode(stream, lo, lo + n, len, code);
        CHECK(got == want, "sf_encode rep %d n %zu maxlen %d", rep, n, maxlen);
        Bytes got2 = emu_sf_encode(stream, lo, lo + n, len, code, 37 + rep);   // tiny staging windows
        CHECK(got2 == want, "sf_encode (windowed) rep %d n %zu maxlen %d", rep, n, maxlen);
        bool bad;
        int passes;
        Bytes dec = emu_sf_decode(want, n, len, code, bad, &passes);
        max_passes = std::max(max_passes, passes);
        CHECK(!bad && dec == raw, "sf_decode rep %d n %zu", rep, n);
        Bytes decf = emu_sf_decode_fused(want, n, len, code, bad);
        CHECK(!bad && decf == raw, "sf_decode_fused rep %d n %zu kind %d maxlen %d", rep, n, kind, maxlen);
    }
    printf("sf codec: max stitch passes %d\n", max_passes);
    // long codes (> 12 bits) through the fallback search, decode only
    {
        uint64_t hist[256] = {0};
        uint64_t a = 1, b = 1;
        for (int i = 0; i < 30; i++) {
            hist[i * 3] = a;
            uint64_t c = a + b;
            a = b;
            b = c;
        }
        uint8_t len[256];
        uint64_t code[256];
        const int maxlen = orc_sf_codes_packed(hist, len, code);
        static char codes[256][ORC_MAX_CODE];
        codes_to_strings(len, code, codes);
        Bytes raw;
        for (int i = 0; i < 30; i++) raw.insert(raw.end(), std::min<uint64_t>(hist[i * 3], 3000), (uint8_t)(i * 3));
        std::shuffle(raw.begin(), raw.end(), rng);
        Bytes enc((raw.size() * maxlen + 7) / 8 + 8);
        enc.resize(orc_sf_encode_block(raw.data(), raw.size(), codes, enc.data()));
        bool bad;
        Bytes dec = emu_sf_decode(enc, raw.size(), len, code, bad, nullptr);
        CHECK(!bad && dec == raw, "sf_decode long codes (max %d)", maxlen);
        Bytes decf = emu_sf_decode_fused(enc, raw.size(), len, code, bad);
        CHECK(!bad && decf == raw, "sf_decode_fused long codes (max %d)", maxlen);
    }
    // larger inputs for the fused decoder: many sequences, skewed and near-uniform alphabets, two symbols (1-bit codes:
    // the record lists overflow and the direct path runs)
    for (int rep = 0; rep < 10; rep++) {
        size_t n = 200000 + rng() % 300000;
        Bytes raw(n);
        if (rep < 4) {
            raw = gen(rng, n, rep);
        } else if (rep < 6) {
            for (auto &v : raw) v = (uint8_t)(rng() & 1);
        } else if (rep < 8) {
            for (auto &v : raw) v = (uint8_t)((rng() % 100) < 90 ? 7 : (rng() % 5));
        } else {
            for (auto &v : raw) v = (uint8_t)(rng() % 200);
        }
        uint64_t hist[256];
        orc_histogram(raw.data(), n, hist);
        int distinct = 0;
        for (int s2 = 0; s2 < 256; s2++) distinct += hist[s2] != 0;
        if (distinct < 2) continue;
        uint8_t len[256];
        uint64_t code[256];
        const int maxlen = orc_sf_codes_packed(hist, len, code);
        static char codes[256][ORC_MAX_CODE];
        codes_to_strings(len, code, codes);
        Bytes enc((n * maxlen + 7) / 8 + 8);
        enc.resize(orc_sf_encode_block(raw.data(), n, codes, enc.data()));
        bool bad;
        Bytes decf = emu_sf_decode_fused(enc, n, len, code, bad);
        CHECK(!bad && decf == raw, "sf_decode_fused big rep %d n %zu maxlen %d", rep, n, maxlen);
        if (getenv("EMU_VERBOSE")) printf("  big rep %d maxlen %d: cumulative met/left/rebuilt %d/%d/%d\n", rep, maxlen, g_fused_kind[0], g_fused_kind[1], g_fused_kind[2]);
    }
    printf("lean multi-code walks checked: %ld\n", g_lean_walks);
    printf("fused decoder: splices met/left/rebuilt %d/%d/%d, speculation failures %d\n", g_fused_kind[0], g_fused_kind[1], g_fused_kind[2],
           g_fused_specfail);
}

int main()
{
    std::mt19937 rng(12345);
    test_rle_encode(rng);
    printf("rle_encode done (%d failures so far)\n", g_fail);
    test_rle_decode(rng);
    printf("rle_decode done (%d failures so far)\n", g_fail);
    test_sf_build(rng);
    printf("sf_build done (%d failures so far)\n", g_fail);
    test_sf_codec(rng);
    printf("sf_codec done (%d failures so far)\n", g_fail);
    if (g_fail) {
        printf("EMU FAILED: %d\n", g_fail);
        return 1;
    }
    printf("EMU OK\n");
    return 0;
}
