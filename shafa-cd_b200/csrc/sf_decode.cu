// sf_decode.cu -- Shannon-Fano prefix decode (reference create_tree +
// shafa_block_decompressor, d.c:422-551), parallel inside a block.
//
// The .shaf format has no sync points, so this is self-synchronising
// sub-sequence decoding (Weissenberger & Schmidt, ICPP'18):
//   k_dec_tables   per block: codes sorted by left-aligned bits, 12-bit LUT (one code),
//                  lut2 (all whole codes of a window: count + bits), lut3 (first two codes)
//   k_dec_sync     persistent CTAs, each owning a contiguous chunk of 4 KiB "sequences";
//                  one thread per 128-bit sub-sequence: decode speculatively from the
//                  sub-sequence start, then restart from where the left neighbour
//                  really ended until nothing changes (later rounds are compacted).
//                  A sequence's first sub-sequence starts where the chunk's previous
//                  sequence ended; per-sequence symbol counts are produced here
//   k_dec_classes  sequences of "lean" blocks (complete code of nearly one length, <= 12 bits, >= 256 KiB of
//                  payload) and of all other blocks, each class numbered densely on the device
//   k_dec_lut16 +  the lean blocks' synchronisation: such codes never re-synchronise, so every start is
//   k_dec_sync_lean walked -- with a 16-bit window table holding ALL whole codes of the window, because
//                  that walk is bound by shared-memory table gathers (see the kernel's comment)
//   k_dec_stitch   across chunk boundaries: feed each sequence the true end of its
//                  predecessor and re-walk until it meets the recorded state;
//                  repeated while a sequence end moved
//   k_dec_scan     symbols per sequence -> output offsets
//   k_dec_write    final decode from the agreed start points (two codes per lookup),
//                  symbols staged and written at their scanned offsets, clipped at
//                  the block's symbol count
// Algorithmic bytes: Z (payload in) + S (symbols out).
#include "sf_decode.cuh"

namespace shafa {

// ---- tables -----------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_dec_tables(const uint8_t *__restrict__ lens, const uint64_t *__restrict__ codes, DecTablesGlobal g, int force_exact)
{
    __shared__ uint64_t left[256];
    __shared__ uint8_t len_s[256];
    __shared__ uint64_t sl[256];
    __shared__ uint8_t sn[256];
    __shared__ uint8_t ss[256];
    __shared__ uint16_t lut1[DEC_LUT_SIZE];
    __shared__ uint32_t len_count[65];
    __shared__ uint32_t s_flags;
    const unsigned s = threadIdx.x;
    const uint64_t b = blockIdx.x;
    const uint32_t l = lens[b * 256 + s];
    const uint64_t c = codes[b * 256 + s];
    const uint64_t lf = l ? (c << (64 - l)) : ~0ull;
    left[s] = lf;
    len_s[s] = (uint8_t)l;
    if (s < 65) len_count[s] = 0;
    __syncthreads();
    if (l && l <= 64) atomicAdd(&len_count[l], 1u);
    // order by (left-aligned code, absent last, symbol)
    uint32_t rank = 0;
    for (int t = 0; t < 256; t++) {
        const uint64_t o = left[t];
        const uint32_t ol = len_s[t];
        const bool before = (o < lf) || (o == lf && ((ol != 0 && l == 0) || ((ol != 0) == (l != 0) && (unsigned)t < s)));
        rank += before;
    }
    sl[rank] = lf;
    sn[rank] = (uint8_t)l;
    ss[rank] = (uint8_t)s;
    __syncthreads();
    g.sleft[b * 256 + s] = sl[s];
    g.slen[b * 256 + s] = sn[s];
    g.ssym[b * 256 + s] = ss[s];
    // ---- what kind of code is this (thread r looks at rank r) ----
    {
        const uint32_t lr = sn[s];
        // prefix-free <=> no code is a prefix of its successor in sorted order (equal codes included)
        const bool clash = lr != 0 && s + 1 < 256 && sn[s + 1] != 0 && ((sl[s + 1] >> (64 - lr)) == (sl[s] >> (64 - lr)));
        const int any_clash = __syncthreads_or(clash);
        if (s == 0) {
            // Kraft sum in units of 2^-64 as a 128-bit number: complete <=> it is exactly 2^64
            uint64_t lo = 0, hi = 0;
            uint32_t minlen = 0, maxlen = 0;
            for (uint32_t k = 1; k <= 64; k++) {
                const uint64_t cnt = len_count[k];
                if (!cnt) continue;
                if (!minlen) minlen = k;
                maxlen = k;
                const uint64_t add_lo = cnt << (64 - k);
                const uint64_t add_hi = (k == 64) ? 0ull : (cnt >> k);
                const uint64_t nlo = lo + add_lo;
                hi += add_hi + (nlo < lo ? 1ull : 0ull);
                lo = nlo;
            }
            uint32_t f = 0;
            if (!any_clash && hi == 1ull && lo == 0ull) f |= DECB_COMPLETE;
            if (2u * minlen > maxlen) f |= DECB_NOSYNC;
            if (maxlen > 32u) f |= DECB_WIDE;
            if (maxlen <= (uint32_t)DEC_LUT_BITS) f |= DECB_SHORT;
            s_flags = f;
            g.flags[b] = (uint8_t)f;
        }
        __syncthreads();
    }
    const bool fused = !force_exact && dec_block_fused_ok((uint8_t)s_flags);
    for (int e = s; e < DEC_LUT_SIZE; e += 256) {
        const uint64_t window = (uint64_t)e << (64 - DEC_LUT_BITS);
        int lo = 0, hi = 256;
        while (lo < hi) {
            const int m = (lo + hi) >> 1;
            if (sl[m] <= window)
                lo = m + 1;
            else
                hi = m;
        }
        uint16_t entry = 0;
        if (lo > 0) {
            const int r = lo - 1;
            const uint32_t rl = sn[r];
            if (rl != 0 && rl <= (uint32_t)DEC_LUT_BITS && (window >> (64 - rl)) == (sl[r] >> (64 - rl)))
                entry = (uint16_t)(((uint32_t)ss[r] << 8) | rl);
            else
                entry = (uint16_t)DEC_LONG;
        }
        lut1[e] = entry;
        if (!fused) g.lut[b * DEC_LUT_SIZE + e] = entry;
    }
    __syncthreads();
    // only the tables of the path the block takes are written (6,554 blocks x 50 KB otherwise)
    if (fused) {
        for (int e = s; e < DEC_LUT_SIZE; e += 256) g.lutp[b * DEC_LUT_SIZE + e] = decf_entry(lut1, (uint32_t)e);
    } else {
        for (int e = s; e < DEC_LUT_SIZE; e += 256) {
            g.lut2[b * DEC_LUT2_SIZE + 2 * e] = dec_lut2_entry(lut1, 2u * (uint32_t)e);
            g.lut2[b * DEC_LUT2_SIZE + 2 * e + 1] = dec_lut2_entry(lut1, 2u * (uint32_t)e + 1u);
            g.lut3[b * DEC_LUT_SIZE + e] = dec_lut3_entry(lut1, (uint32_t)e);
        }
    }
}

// per sequence of the fused plan: where its payload lies and how much of it to copy (one CTA per block)
__global__ void __launch_bounds__(256)
k_dec_seqrec(DecPlan plan, DecSeqRec *__restrict__ rec)
{
    const uint32_t b = blockIdx.x;
    const uint64_t s0 = plan.seq0[b], s1 = plan.seq0[b + 1];
    const uint64_t slot = (plan.z[b] + 15) & ~15ull, off0 = plan.pay_off[b];
    for (uint64_t i = threadIdx.x; s0 + i < s1; i += 256) {
        const uint64_t off = off0 + i * DEC_SEQ_BYTES;
        const uint64_t left_bytes = slot - i * DEC_SEQ_BYTES;
        const uint32_t bytes = left_bytes < (uint64_t)DEC_SEQ_BYTES + 16 ? (uint32_t)left_bytes : (uint32_t)DEC_SEQ_BYTES + 16u;
        rec[s0 + i] = make_uint4((uint32_t)off, (uint32_t)(off >> 32), b, bytes);
    }
}

// ---- shared helpers -----------------------------------------------------------------------
constexpr int DEC_MAP_MAX = 16;     // longest code for which the all-starts fallback of k_dec_sync is used
constexpr int DEC_MAX_ROUNDS = 12;  // sync rounds before that fallback is taken in any case
constexpr int DEC_STAGE = 12288;  // symbols of one sequence staged in shared memory before the coalesced copy-out

struct DecShared {
    uint64_t pairs[DEC_PAIR_SLOTS];  // row of sub-sequence s at [s * DEC_ROW, (s + 1) * DEC_ROW) (core.h)
    uint16_t lut1[DEC_LUT_SIZE];
    uint64_t sleft[256];
    uint8_t slen[256];
    uint8_t ssym[256];
    uint32_t scan[DEC_THREADS / 32 + 1];
};

// lut2_dst: shared-memory copy of the multi-symbol count table (sync kernel only), or nullptr
__device__ __forceinline__ void dec_load_tables(DecShared &sh, const DecTablesGlobal &g, uint32_t b, uint8_t *lut2_dst)
{
    const unsigned tid = threadIdx.x;
    const uint4 *src1 = reinterpret_cast<const uint4 *>(g.lut + (uint64_t)b * DEC_LUT_SIZE);
    const uint4 *src2 = reinterpret_cast<const uint4 *>(g.lut2 + (uint64_t)b * DEC_LUT2_SIZE);  // 8 KiB, like lut1
    uint4 *dst1 = reinterpret_cast<uint4 *>(sh.lut1);
    uint4 *dst2 = reinterpret_cast<uint4 *>(lut2_dst);
    for (int i = tid; i < DEC_LUT_SIZE * 2 / 16; i += DEC_THREADS) {
        dst1[i] = src1[i];
        if (lut2_dst) dst2[i] = src2[i];
    }
    sh.sleft[tid] = g.sleft[(uint64_t)b * 256 + tid];
    sh.slen[tid] = g.slen[(uint64_t)b * 256 + tid];
    sh.ssym[tid] = g.ssym[(uint64_t)b * 256 + tid];
}

// The loads are issued one sequence ahead (dec_fetch_row) and only turned into a row when that sequence's turn comes.
__device__ __forceinline__ RowRegs dec_fetch_row(const uint8_t *payload, uint64_t pay_off, uint64_t z, uint64_t seq_in_block)
{
    const uint64_t o = seq_in_block * DEC_SEQ_BYTES + (uint64_t)threadIdx.x * 16;  // offset in the block's payload
    const uint64_t slot = (z + 15) & ~15ull;                                       // readable bytes of the block's slot
    const uint8_t *base = payload + pay_off;
    RowRegs r;
    r.v = make_uint4(0, 0, 0, 0);
    r.u = make_uint2(0, 0);
    if (o < slot) r.v = __ldg(reinterpret_cast<const uint4 *>(base + o));
    // the first half of the next thread's vector: the same cache line, or the next one
    if (o + 16 < slot) r.u = __ldg(reinterpret_cast<const uint2 *>(base + o + 16));
    return r;
}
// Which block a sequence belongs to, tracked incrementally while a CTA walks its chunk (every thread
// keeps the same copy: no broadcast through shared memory, no barrier).  `seq0` is the numbering the
// kernel walks: all sequences (plan.seq0), or the sequences of one class (plan.cls_seq0[c]).
struct SeqCursor {
    const uint64_t *seq0;
    uint32_t b;
    uint64_t first, next;  // block b holds sequences [first, next) of that numbering
    __device__ __forceinline__ void init(const uint64_t *numbering, uint32_t nb, uint64_t seq)
    {
        seq0 = numbering;
        b = find_block(seq0, nb, seq);
        first = seq0[b];
        next = seq0[b + 1];
    }
    __device__ __forceinline__ void advance_to(uint64_t seq)
    {
        while (seq >= next) {  // skips blocks without sequences of this numbering
            b++;
            first = next;
            next = seq0[b + 1];
        }
    }
};

// what a sequence needs from global memory, fetched one sequence ahead
struct SeqFetch {
    uint32_t b;
    uint64_t sib, z;
    RowRegs row;
};
__device__ __forceinline__ SeqFetch dec_fetch_seq(const uint8_t *payload, const DecPlan &plan, SeqCursor &c, uint64_t seq)
{
    c.advance_to(seq);
    SeqFetch f;
    f.b = c.b;
    f.sib = seq - c.first;  // a block's sequences all belong to one class: the same index in every numbering
    f.z = plan.z[c.b];
    f.row = dec_fetch_row(payload, plan.pay_off[c.b], f.z, f.sib);
    return f;
}

// ---- lean blocks: the 16-bit multi-code table and the two sequence classes ---------------------------------------
__global__ void __launch_bounds__(256)
k_dec_lut16(DecPlan plan, DecTablesGlobal g)
{
    // lut1 with one pad entry after every 16: the windows of a warp's second codes are 16 entries apart (the low four
    // bits of a 12-bit window past the end of a 16-bit one are zero) and would all fall into one bank
    __shared__ uint16_t lut1p[DEC_LUT_SIZE + DEC_LUT_SIZE / 16];
    const uint32_t b = blockIdx.x;
    if (!dec_block_lean(g.flags[b], plan.z[b])) return;
    for (int e = threadIdx.x; e < DEC_LUT_SIZE; e += 256) lut1p[e + (e >> 4)] = g.lut[(uint64_t)b * DEC_LUT_SIZE + e];
    __syncthreads();
    uint32_t *dst = reinterpret_cast<uint32_t *>(g.lut16 + (uint64_t)plan.lean_slot[b] * DEC_LUT16_SIZE);
    for (uint32_t q = threadIdx.x; q < (uint32_t)DEC_LUT16_SIZE / 4; q += 256) {
        uint32_t four = 0;
#pragma unroll
        for (int k = 0; k < 4; k++) {
            // core.h: dec_lut16_entry, reading the padded table
            const uint32_t v = 4u * q + (uint32_t)k;
            uint32_t pos = 0, n = 0;
            while (n < 7u) {
                const uint32_t i = ((v << pos) & 0xFFFFu) >> (DEC_LUT16_BITS - DEC_LUT_BITS);
                const uint32_t l = lut1p[i + (i >> 4)] & 0xFFu;
                if (l - 1u >= (uint32_t)DEC_LUT_BITS || pos + l > (uint32_t)DEC_LUT16_BITS) break;
                pos += l;
                n++;
            }
            four |= (n ? ((n << 5) | pos) : ((1u << 5) | 1u)) << (8 * k);
        }
        dst[q] = four;
    }
}

// one CTA: number the sequences of the two classes in block order, hand every lean block its table slot
// out: cls[0][nb+1], cls[1][nb+1] (u64), slot[nb] (u32)
__global__ void __launch_bounds__(256)
k_dec_classes(const uint64_t *__restrict__ seq0, const uint64_t *__restrict__ z, const uint8_t *__restrict__ flags, uint32_t nb,
              uint64_t *__restrict__ cls0, uint64_t *__restrict__ cls1, uint32_t *__restrict__ slot)
{
    __shared__ uint32_t scan[9];
    __shared__ uint64_t carry[3];
    if (threadIdx.x < 3) carry[threadIdx.x] = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += 256) {
        const uint32_t b = base + threadIdx.x;
        uint32_t n0 = 0, n1 = 0, lean = 0;
        if (b < nb) {
            const uint32_t ns = (uint32_t)(seq0[b + 1] - seq0[b]);
            lean = dec_block_lean(flags[b], z[b]) ? 1u : 0u;
            n0 = lean ? 0u : ns;
            n1 = lean ? ns : 0u;
        }
        uint32_t t0, t1, t2;
        const uint32_t e0 = block_excl_scan<256>(n0, scan, &t0);
        __syncthreads();
        const uint32_t e1 = block_excl_scan<256>(n1, scan, &t1);
        __syncthreads();
        const uint32_t e2 = block_excl_scan<256>(lean, scan, &t2);
        if (b < nb) {
            cls0[b] = carry[0] + e0;
            cls1[b] = carry[1] + e1;
            slot[b] = (uint32_t)carry[2] + e2;
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            carry[0] += t0;
            carry[1] += t1;
            carry[2] += t2;
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        cls0[nb] = carry[0];
        cls1[nb] = carry[1];
    }
}

// The decode loops read the tables through 32-bit shared-memory addresses computed once per sequence
// (explicit ld.shared), so that no generic-to-shared address arithmetic is left inside them.
struct DecViewS {
    uint32_t pairs_s, lut1_s, lut2_s, lut3_s;  // shared-window addresses; 0 = table not loaded
    const uint16_t *lut1;                      // the same tables as pointers for the rare long-code search
    const uint64_t *sleft;
    const uint8_t *slen;
    const uint8_t *ssym;
    __device__ __forceinline__ uint64_t pair(uint32_t slot) const
    {
        uint64_t v;
        asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(pairs_s + slot * 8u));
        return v;
    }
    __device__ __forceinline__ uint32_t l1(uint32_t i) const
    {
        uint32_t v;
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(lut1_s + i * 2u));
        return v;
    }
    __device__ __forceinline__ uint32_t l1len(uint32_t i) const  // low byte of the entry: the code length
    {
        uint32_t v;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(lut1_s + i * 2u));
        return v;
    }
    __device__ __forceinline__ uint32_t l2(uint32_t i) const
    {
        uint32_t v;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(lut2_s + i));
        return v;
    }
    __device__ __forceinline__ uint32_t l3(uint32_t i) const
    {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(lut3_s + i * 4u));
        return v;
    }
    __device__ __forceinline__ bool has_l3() const { return lut3_s != 0; }
};

__device__ __forceinline__ DecViewS dec_view_of(const DecShared &sh, const uint8_t *lut2, const uint32_t *lut3)
{
    DecViewS t;
    t.pairs_s = smem_u32(sh.pairs);
    t.lut1_s = smem_u32(sh.lut1);
    t.lut2_s = lut2 ? smem_u32(lut2) : 0u;
    t.lut3_s = lut3 ? smem_u32(lut3) : 0u;
    // opaque to the optimiser: otherwise the addresses are rebuilt from the CTA's shared window inside the loops
    asm volatile("" : "+r"(t.pairs_s), "+r"(t.lut1_s), "+r"(t.lut2_s), "+r"(t.lut3_s));
    t.lut1 = sh.lut1;
    t.sleft = sh.sleft;
    t.slen = sh.slen;
    t.ssym = sh.ssym;
    return t;
}

// sub-sequence record: bits [6:0] start offset inside the sub-sequence (< 64), bits [15:7] symbol count (<= 128)
__device__ __forceinline__ uint16_t sub_pack(uint32_t start, uint32_t count) { return (uint16_t)(start | (count << 7)); }
__device__ __forceinline__ uint32_t sub_start(uint16_t v) { return v & 127u; }
__device__ __forceinline__ uint32_t sub_count(uint16_t v) { return v >> 7; }

// contiguous chunk of sequences owned by this CTA (tables are reloaded only when the block changes)
__device__ __forceinline__ void dec_chunk(uint64_t nseq, uint64_t *first, uint64_t *last)
{
    const uint64_t per = (nseq + gridDim.x - 1) / gridDim.x;
    *first = (uint64_t)blockIdx.x * per;
    const uint64_t e = *first + per;
    *last = e < nseq ? e : nseq;
}

// ---- phase 1: synchronise inside each sequence ----------------------------------------------
__global__ void __launch_bounds__(DEC_THREADS, 6)
k_dec_sync(const uint8_t *__restrict__ payload, DecPlan plan, DecTablesGlobal g, uint16_t *__restrict__ sub_info,
           uint32_t *__restrict__ seq_end, uint32_t *__restrict__ seq_count)
{
    // this kernel's sequences: class 0 (everything but the lean blocks), numbered densely by k_dec_classes
    const uint64_t nseq = plan.cls_seq0[0][plan.nb];
    __shared__ DecShared sh;
    __shared__ __align__(16) uint8_t s_lut2[DEC_LUT2_SIZE];
    __shared__ uint32_t s_ends[DEC_THREADS];    // overshoot of sub-sequence i into sub-sequence i+1
    __shared__ uint16_t s_start[DEC_THREADS];   // start offset the current record of sub-sequence i was decoded from
    __shared__ uint16_t s_count[DEC_THREADS];
    __shared__ uint32_t s_list[DEC_THREADS];    // compacted work list of a sparse round: sub-sequence | start << 16
    __shared__ uint32_t s_wcnt[DEC_THREADS / 32];
    __shared__ uint32_t s_nlist;
    __shared__ uint32_t s_wsum[DEC_THREADS / 32];
    // Tables of the all-starts path.  They alias shared memory that path does not need: the maps take the place of lut2
    // (all starts are walked with lut1), the chain paths the place of the payload rows (every thread has moved its row into
    // registers by then).  They were in a global scratch slice before: on C4 (every block takes this path) the 2-byte map
    // stores of a warp, 32 bytes apart, cost 18 x the payload in L2 sector writes (dec_sync 8.7 -> 6.7 ms/GiB).
    static_assert(DEC_THREADS * DEC_MAP_MAX * 2 <= DEC_LUT2_SIZE, "maps fit the lut2 area");
    static_assert((DEC_THREADS / 32) * DEC_MAP_MAX * 33 <= DEC_PAIR_SLOTS * 8, "paths fit the row area");
    uint16_t *const s_map = reinterpret_cast<uint16_t *>(s_lut2);                             // [sub-sequence][start] -> (overshoot << 10) | count
    uint8_t *const s_path = reinterpret_cast<uint8_t *>(sh.pairs);                            // [warp][entry offset][i]: start of the warp's i-th sub-sequence
    uint8_t *const s_path_out = s_path + (DEC_THREADS / 32) * DEC_MAP_MAX * 32;               // [warp][entry offset]: offset handed to the next warp
    __shared__ uint32_t s_maxlen;
    __shared__ uint32_t s_nosync;  // the block's code has nearly one length (k_dec_tables): go to all starts at once
    __shared__ uint32_t s_lean;    // ... and it is complete with no code longer than the lut1 window: lengths come straight from lut1
    const unsigned tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    uint64_t seq_first, seq_last;
    dec_chunk(nseq, &seq_first, &seq_last);
    uint32_t cur_block = 0xffffffffu;
    uint32_t carry = 0;  // where the previous sequence of this CTA's chunk left its last sub-sequence
    if (seq_first >= seq_last) return;
    SeqCursor cursor;
    cursor.init(plan.cls_seq0[0], plan.nb, seq_first);
    SeqFetch cur = dec_fetch_seq(payload, plan, cursor, seq_first);
    for (uint64_t cseq = seq_first; cseq < seq_last; cseq++) {
        SeqFetch nxt = cur;
        if (cseq + 1 < seq_last) nxt = dec_fetch_seq(payload, plan, cursor, cseq + 1);  // in flight during this sequence
        const uint32_t b = cur.b;
        const uint64_t sib = cur.sib, z = cur.z;
        const uint64_t seq = plan.seq0[b] + sib;  // its number among ALL sequences: where its records go
        __syncthreads();  // previous sequence fully consumed
        if (b != cur_block) {
            cur_block = b;
            dec_load_tables(sh, g, b, s_lut2);
            if (tid == 0) {
                s_maxlen = 0;
                s_nosync = (g.flags[b] & DECB_NOSYNC) != 0;
                s_lean = (g.flags[b] & DECB_COMPLETE) != 0;
            }
            __syncthreads();
            const uint32_t wmax = __reduce_max_sync(FULL, (uint32_t)sh.slen[tid]);
            if (lane == 0) atomicMax(&s_maxlen, wmax);
            {
                // the lean all-starts loop trusts lut1 blindly (a zero length would never advance): look at every entry once
                uint32_t odd = 0;
                const uint32_t *l32 = reinterpret_cast<const uint32_t *>(sh.lut1);
                for (int i = tid; i < DEC_LUT_SIZE / 2; i += DEC_THREADS) {
                    const uint32_t e2 = l32[i];
                    odd |= (uint32_t)(((e2 & 0xFFu) - 1u >= (uint32_t)DEC_LUT_BITS) | (((e2 >> 16) & 0xFFu) - 1u >= (uint32_t)DEC_LUT_BITS));
                }
                if (__syncthreads_or((int)odd) && tid == 0) s_lean = 0;
            }
        }
        // The first sub-sequence starts where the previous sequence ended.  Inside a chunk that is known;
        // the first sequence of a chunk assumes 0 and k_dec_stitch repairs it (a block's first sequence IS 0).
        if (sib == 0 || cseq == seq_first) carry = 0;
        dec_store_row(sh.pairs, cur.row, z, sib);
        cur = nxt;
        __syncthreads();
        const DecViewS t = dec_view_of(sh, s_lut2, nullptr);
        // bits of this sequence that exist in the payload
        const uint64_t bits_left = (z - sib * DEC_SEQ_BYTES) * 8;
        const uint32_t limit = bits_left < (uint64_t)DEC_SEQ_BYTES * 8 ? (uint32_t)bits_left : DEC_SEQ_BYTES * 8;

        // decode sub-sequence `sid` from offset `st` and record where it leaves it
        auto run = [&](uint32_t sid, uint32_t st) {
            const uint32_t sub_lo = sid * DEC_SUB_BITS;
            const uint32_t stop = limit > sub_lo ? (limit - sub_lo < (uint32_t)DEC_SUB_BITS ? limit - sub_lo : (uint32_t)DEC_SUB_BITS) : 0u;
            const DecSub r = dec_count_subsequence(t, sid * DEC_ROW, st, stop);
            s_start[sid] = (uint16_t)st;
            s_count[sid] = (uint16_t)r.count;
            s_ends[sid] = r.end > (uint32_t)DEC_SUB_BITS ? r.end - DEC_SUB_BITS : 0;  // overshoot into the next sub-sequence
        };
        const uint32_t maxlen = s_maxlen;  // (written before the barrier above)
        // codes of nearly one length: the speculative rounds would be wasted, take every start right away
        const bool all_starts_now = s_nosync && maxlen <= (uint32_t)DEC_MAP_MAX;
        if (all_starts_now)
            s_start[tid] = 0xffffu;  // "not decoded yet": every sub-sequence counts as changed below
        else
            run(tid, tid ? 0u : carry);
        uint32_t round = 1;
        // Rounds until every sub-sequence starts where its predecessor ended.  Round 2 moves almost every
        // start; later rounds touch only the few sub-sequences whose predecessor had not synchronised yet,
        // so their work is compacted into the first warps instead of costing every warp a full pass.
        for (;;) {
            __syncthreads();
            const uint32_t want = tid ? s_ends[tid - 1] : carry;
            const bool changed = (want != s_start[tid]);
            if (tid == 0) s_nlist = 0;
            const uint32_t total = (uint32_t)__syncthreads_count(changed);
            if (total == 0) break;
            round++;
            if ((all_starts_now || (round > 3 && total > DEC_THREADS / 2) || round > DEC_MAX_ROUNDS) && maxlen <= (uint32_t)DEC_MAP_MAX) {
                // Codes of (nearly) one length never re-synchronise, and the rounds above would then settle
                // one sub-sequence at a time.  Instead every thread decodes its sub-sequence from EVERY start
                // that can occur (an overshoot is shorter than the longest code) and one thread follows the
                // chain through the resulting tables.
                {
                    const uint32_t sub_lo = tid * DEC_SUB_BITS;
                    const uint32_t stop = limit > sub_lo ? (limit - sub_lo < (uint32_t)DEC_SUB_BITS ? limit - sub_lo : (uint32_t)DEC_SUB_BITS) : 0u;
                    // The sub-sequence's six words live in registers for all `maxlen` walks: a window is one funnel shift of
                    // two of them (the word index is a compile-time loop counter), not a shared-memory load per code.
                    uint32_t w[DEC_ROW_WORDS];
#pragma unroll
                    for (int k = 0; k < DEC_ROW; k++) {
                        const uint64_t pr = t.pair(tid * DEC_ROW + k);
                        w[k] = (uint32_t)(pr >> 32);
                        if (k == DEC_ROW - 1) w[k + 1] = (uint32_t)pr;
                    }
                    if (s_lean) {
                        // a complete code without long codes: every 12-bit window holds a code, its length is the low
                        // byte of the lut1 entry -- one funnel shift, one byte load, two adds per code and nothing else
                        for (uint32_t st = 0; st < maxlen; st++) {
                            uint32_t q = st, count = 0;
#pragma unroll
                            for (int i = 0; i < DEC_SUB_BITS / 32; i++) {
                                const uint32_t lim = stop < 32u * (i + 1) ? stop : 32u * (i + 1);
                                while (q < lim) {
                                    q += t.l1len(__funnelshift_l(w[i + 1], w[i], q) >> (32 - DEC_LUT_BITS));
                                    count++;
                                }
                            }
                            const uint32_t over = q > (uint32_t)DEC_SUB_BITS ? q - DEC_SUB_BITS : 0u;
                            s_map[tid * DEC_MAP_MAX + st] = (uint16_t)((over << 10) | count);
                        }
                    } else {
                        for (uint32_t st = 0; st < maxlen; st++) {
                            uint32_t q = st, count = 0, bad_dummy = 0;
#pragma unroll
                            for (int i = 0; i < DEC_SUB_BITS / 32; i++) {
                                const uint32_t lim = stop < 32u * (i + 1) ? stop : 32u * (i + 1);
                                while (q < lim) {
                                    const uint32_t e = t.l1(__funnelshift_l(w[i + 1], w[i], q) >> (32 - DEC_LUT_BITS));
                                    uint32_t l = e & 0xFFu;
                                    if (l - 1u >= (uint32_t)DEC_LUT_BITS) {  // a code longer than the table window (or none): generic step
                                        uint32_t sym;
                                        l = dec_one(t, tid * DEC_ROW, q, &sym, &bad_dummy);
                                    }
                                    q += l;
                                    count++;
                                }
                            }
                            const uint32_t over = q > (uint32_t)DEC_SUB_BITS ? q - DEC_SUB_BITS : 0u;
                            s_map[tid * DEC_MAP_MAX + st] = (uint16_t)((over << 10) | count);
                        }
                    }
                }
                __syncthreads();
                // the chain, two levels: lane s of every warp follows its warp's 32 sub-sequences from entry
                // offset s (all offsets and warps in parallel), one thread links the eight warps, and every
                // thread then picks the path that belongs to its warp's real entry offset
                if (lane < maxlen) {
                    uint32_t st = lane;
                    for (int i = 0; i < 32; i++) {
                        s_path[(warp * DEC_MAP_MAX + lane) * 32 + i] = (uint8_t)st;
                        st = s_map[(warp * 32 + i) * DEC_MAP_MAX + st] >> 10;  // < maxlen: shorter than the longest code
                    }
                    s_path_out[warp * DEC_MAP_MAX + lane] = (uint8_t)st;
                }
                __syncthreads();
                if (tid == 0) {
                    uint32_t st = carry < maxlen ? carry : 0u;  // carry >= maxlen cannot happen
                    for (int w = 0; w < DEC_THREADS / 32; w++) {
                        s_wcnt[w] = st;
                        st = s_path_out[w * DEC_MAP_MAX + st];
                    }
                }
                __syncthreads();
                {
                    const uint32_t st = s_path[(warp * DEC_MAP_MAX + s_wcnt[warp]) * 32 + lane];
                    const uint32_t e = s_map[tid * DEC_MAP_MAX + st];
                    s_start[tid] = (uint16_t)st;
                    s_count[tid] = (uint16_t)(e & 1023u);
                    s_ends[tid] = e >> 10;
                }
                __syncthreads();
                if (!all_starts_now) cur_block = 0xffffffffu;  // the maps overwrote lut2: it has to come back before the next sequence
                break;
            }
            if (total > DEC_THREADS / 4) {
                if (changed) run(tid, want);
            } else {
                // work list in any order: one shared counter, one atomic per warp
                const unsigned bal = __ballot_sync(FULL, changed);
                uint32_t base = 0;
                if (lane == 0 && bal) base = atomicAdd(&s_nlist, (uint32_t)__popc(bal));
                base = __shfl_sync(FULL, base, 0);
                if (changed) s_list[base + (uint32_t)__popc(bal & ((1u << lane) - 1u))] = tid | (want << 16);
                __syncthreads();
                if (tid < total) {
                    const uint32_t e = s_list[tid];
                    run(e & 0xffffu, e >> 16);
                }
            }
        }
        const uint32_t my_count = s_count[tid];
        sub_info[seq * DEC_THREADS + tid] = sub_pack(s_start[tid], my_count);
        carry = s_ends[DEC_THREADS - 1];
        if (tid == DEC_THREADS - 1) seq_end[seq] = carry;
        // symbols of the sequence (k_dec_stitch adjusts it when it re-decodes a sub-sequence)
        const uint32_t wsum = __reduce_add_sync(FULL, my_count);
        if (lane == 0) s_wsum[warp] = wsum;
        __syncthreads();
        if (tid == 0) {
            uint32_t tot = 0;
#pragma unroll
            for (int w = 0; w < DEC_THREADS / 32; w++) tot += s_wsum[w];
            seq_count[seq] = tot;
        }
        // a speculative start may hit garbage; errors are judged in k_dec_write
    }
}

// ---- phase 1 for lean blocks -------------------------------------------------------------------------------------
// Codes of nearly one length never re-synchronise: every thread walks its sub-sequence from EVERY start that can occur
// (an overshoot is shorter than the longest code) and the chain through the resulting maps is followed warp-parallel.
// The walks are bound by shared-memory table gathers (ncu on C4: LSU wavefronts 95 % of peak, 2.8 wavefronts per load
// from bank conflicts on random indices), so this kernel takes fewer of them: a 16-bit window table holding all whole
// codes of the window (core.h: dec_lut16_entry, dec_lean_walk) -- 64 KiB per block, which is why lean blocks have a
// kernel (and a shared-memory budget) of their own.
constexpr int LEAN_SEQS = 2;                            // sequences a CTA walks at a time: they share the 72 KiB of tables, and
constexpr int LEAN_THREADS = LEAN_SEQS * DEC_THREADS;   // 16 warps per CTA x 2 CTAs per SM hide the table latency (8 warps: 22 ms on C4)
constexpr int LEAN_WARPS = LEAN_THREADS / 32;
constexpr int LEAN_MAP_STRIDE = LEAN_THREADS + 2;
#ifndef SHAFA_LEAN_NS
#define SHAFA_LEAN_NS 3
#endif
constexpr int LEAN_NS = SHAFA_LEAN_NS;                  // chains a thread walks at a time (independent: their table loads overlap)
struct DecLeanShared {
    alignas(16) uint8_t lut16[DEC_LUT16_SIZE];
    alignas(16) uint16_t lut1[DEC_LUT_SIZE];         // single-code steps for the last bits of a sub-sequence
    // Layouts chosen by who touches them together (the kernel is bound by shared-memory wavefronts): a warp stores the
    // maps of 32 consecutive sub-sequences for ONE start, the chain lanes read ONE sub-sequence for all starts
    uint16_t map[DEC_MAP_MAX * LEAN_MAP_STRIDE];     // [start][sub-sequence] -> (overshoot << 10) | count; one bank further per start
    uint8_t path[LEAN_WARPS * 32 * DEC_MAP_MAX];     // [warp][i][entry offset]: start of the warp's i-th sub-sequence
    uint8_t path_out[LEAN_WARPS * DEC_MAP_MAX];      // [warp][entry offset]: offset handed to the next warp
    uint32_t wentry[LEAN_WARPS];
    uint32_t wsum[LEAN_WARPS];
    uint32_t wmax[LEAN_WARPS];
    uint32_t last_over;
};
struct DecViewLean {
    uint32_t lut1_s, lut16_s;
    __device__ __forceinline__ uint32_t l1len(uint32_t i) const
    {
        uint32_t v;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(lut1_s + i * 2u));
        return v;
    }
    __device__ __forceinline__ uint32_t l16(uint32_t i) const
    {
        uint32_t v;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(lut16_s + i));
        return v;
    }
};

__global__ void __launch_bounds__(LEAN_THREADS, 2)
k_dec_sync_lean(const uint8_t *__restrict__ payload, DecPlan plan, DecTablesGlobal g, uint16_t *__restrict__ sub_info,
                uint32_t *__restrict__ seq_end, uint32_t *__restrict__ seq_count)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DecLeanShared &sh = *reinterpret_cast<DecLeanShared *>(smem_raw);
    const unsigned tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    const unsigned half = tid / DEC_THREADS, sid = tid % DEC_THREADS;  // which of the CTA's sequences, which sub-sequence of it
    const uint64_t nseq = plan.cls_seq0[1][plan.nb];
    uint64_t seq_first, seq_last;
    dec_chunk(nseq, &seq_first, &seq_last);
    if (seq_first >= seq_last) return;
    uint32_t cur_block = 0xffffffffu, carry = 0, maxlen = 0;
    SeqCursor cursor;
    cursor.init(plan.cls_seq0[1], plan.nb, seq_first);
    DecViewLean t;
    t.lut1_s = smem_u32(sh.lut1);
    t.lut16_s = smem_u32(sh.lut16);
    asm volatile("" : "+r"(t.lut1_s), "+r"(t.lut16_s));  // opaque: not rebuilt from the shared window inside the loops
    uint64_t cseq = seq_first;
    while (cseq < seq_last) {
        cursor.advance_to(cseq);
        const uint32_t b = cursor.b;
        // this round: up to LEAN_SEQS consecutive sequences of the same block (and of this CTA's chunk)
        uint64_t room = (cursor.next < seq_last ? cursor.next : seq_last) - cseq;
        const uint32_t ns = room < (uint64_t)LEAN_SEQS ? (uint32_t)room : (uint32_t)LEAN_SEQS;
        const uint64_t sib0 = cseq - cursor.first, z = plan.z[b];
        const uint64_t sib = sib0 + half;           // my sequence inside the block
        const uint64_t seq = plan.seq0[b] + sib;    // its number among ALL sequences: where its records go
        const bool mine = half < ns;
        RowRegs row;
        row.v = make_uint4(0, 0, 0, 0);
        row.u = make_uint2(0, 0);
        {
            // (dec_fetch_row indexes by threadIdx.x: here the sub-sequence number is sid)
            const uint64_t o = sib * DEC_SEQ_BYTES + (uint64_t)sid * 16;
            const uint64_t slot = (z + 15) & ~15ull;
            const uint8_t *base = payload + plan.pay_off[b];
            if (mine && o < slot) row.v = __ldg(reinterpret_cast<const uint4 *>(base + o));
            if (mine && o + 16 < slot) row.u = __ldg(reinterpret_cast<const uint2 *>(base + o + 16));
        }
        __syncthreads();  // previous round fully consumed
        if (b != cur_block) {
            cur_block = b;
            const uint4 *src16 = reinterpret_cast<const uint4 *>(g.lut16 + (uint64_t)plan.lean_slot[b] * DEC_LUT16_SIZE);
            uint4 *dst16 = reinterpret_cast<uint4 *>(sh.lut16);
#pragma unroll 4
            for (int i = tid; i < DEC_LUT16_SIZE / 16; i += LEAN_THREADS) dst16[i] = src16[i];
            const uint4 *src1 = reinterpret_cast<const uint4 *>(g.lut + (uint64_t)b * DEC_LUT_SIZE);
            uint4 *dst1 = reinterpret_cast<uint4 *>(sh.lut1);
            for (int i = tid; i < DEC_LUT_SIZE * 2 / 16; i += LEAN_THREADS) dst1[i] = src1[i];
            const uint32_t wm = __reduce_max_sync(FULL, tid < 256 ? (uint32_t)g.slen[(uint64_t)b * 256 + tid] : 0u);
            if (lane == 0) sh.wmax[warp] = wm;
            __syncthreads();
            maxlen = 0;
#pragma unroll
            for (int k = 0; k < LEAN_WARPS; k++) maxlen = sh.wmax[k] > maxlen ? sh.wmax[k] : maxlen;
        }
        if (sib0 == 0 || cseq == seq_first) carry = 0;
        uint32_t w[DEC_ROW_WORDS];
        {
            // dec_row_words for sub-sequence sid of sequence sib: bytes at or past z read as zero, words big-endian
            const uint64_t o = sib * DEC_SEQ_BYTES + (uint64_t)sid * 16;
            w[0] = row.v.x, w[1] = row.v.y, w[2] = row.v.z, w[3] = row.v.w, w[4] = row.u.x, w[5] = row.u.y;
            if (o + 4 * DEC_ROW_WORDS > z) {
#pragma unroll
                for (int k = 0; k < DEC_ROW_WORDS; k++) {
                    const uint64_t bo = o + 4 * k;
                    if (bo >= z)
                        w[k] = 0;
                    else if (bo + 4 > z)
                        w[k] &= (0xffffffffu >> (8 * (uint32_t)(bo + 4 - z)));
                }
            }
#pragma unroll
            for (int k = 0; k < DEC_ROW_WORDS; k++) w[k] = bswap32(w[k]);
        }
        __syncthreads();  // tables in place
        uint32_t stop = 0;
        if (mine) {
            const uint64_t bits_left = (z - sib * DEC_SEQ_BYTES) * 8;
            const uint32_t limit = bits_left < (uint64_t)DEC_SEQ_BYTES * 8 ? (uint32_t)bits_left : DEC_SEQ_BYTES * 8;
            const uint32_t sub_lo = sid * DEC_SUB_BITS;
            stop = limit > sub_lo ? (limit - sub_lo < (uint32_t)DEC_SUB_BITS ? limit - sub_lo : (uint32_t)DEC_SUB_BITS) : 0u;
        }
        for (uint32_t st = 0; st < maxlen; st += LEAN_NS) {  // maxlen <= 12: a lean block has no longer code
            uint32_t cnt[LEAN_NS], end[LEAN_NS];
            dec_lean_walk<LEAN_NS>(t, w, st, stop, cnt, end);
#pragma unroll
            for (int k = 0; k < LEAN_NS; k++) {  // (starts at or past maxlen are computed too and never looked at: the map has 16 slots)
                const uint32_t ov = end[k] > (uint32_t)DEC_SUB_BITS ? end[k] - DEC_SUB_BITS : 0u;
                sh.map[(st + k) * LEAN_MAP_STRIDE + tid] = (uint16_t)((ov << 10) | cnt[k]);
            }
        }
        __syncthreads();
        // the chain, two levels: lane s of every warp follows its warp's 32 sub-sequences from entry offset s (all
        // offsets and warps in parallel), one thread links the warps (the round's sequences are consecutive: the chain
        // runs straight through them), and every thread then picks the path that belongs to its warp's real entry offset
        if (lane < maxlen) {
            uint32_t st = lane;
            for (int i = 0; i < 32; i++) {
                sh.path[(warp * 32 + i) * DEC_MAP_MAX + lane] = (uint8_t)st;
                st = sh.map[st * LEAN_MAP_STRIDE + warp * 32 + i] >> 10;  // < maxlen: shorter than the longest code
            }
            sh.path_out[warp * DEC_MAP_MAX + lane] = (uint8_t)st;
        }
        __syncthreads();
        if (tid == 0) {
            uint32_t st = carry < maxlen ? carry : 0u;  // carry >= maxlen cannot happen
            for (int k = 0; k < LEAN_WARPS; k++) {
                sh.wentry[k] = st;
                st = sh.path_out[k * DEC_MAP_MAX + st];
            }
        }
        __syncthreads();
        const uint32_t st = sh.path[(warp * 32 + lane) * DEC_MAP_MAX + sh.wentry[warp]];
        const uint32_t e = sh.map[st * LEAN_MAP_STRIDE + tid];
        const uint32_t my_count = e & 1023u;
        if (mine) sub_info[seq * DEC_THREADS + sid] = sub_pack(st, my_count);
        if (mine && sid == DEC_THREADS - 1) {
            if (half == ns - 1) sh.last_over = e >> 10;
            seq_end[seq] = e >> 10;
        }
        const uint32_t wsum = __reduce_add_sync(FULL, my_count);
        if (lane == 0) sh.wsum[warp] = wsum;
        __syncthreads();
        carry = sh.last_over;
        if (mine && sid == 0) {
            uint32_t tot = 0;
#pragma unroll
            for (int k = 0; k < DEC_THREADS / 32; k++) tot += sh.wsum[half * (DEC_THREADS / 32) + k];
            seq_count[seq] = tot;
        }
        cseq += ns;
    }
}

// ---- phase 2: stitch sequences together -------------------------------------------------------
// one thread per sequence; reads payload bytes straight from global memory (only a
// few sub-sequences per sequence are ever touched)
__device__ __forceinline__ uint64_t window_global(const uint8_t *base, uint64_t z, uint64_t bit)
{
    const uint64_t byte = bit >> 3;
    uint64_t hi = 0;
    uint32_t lo = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) hi = (hi << 8) | (byte + k < z ? base[byte + k] : 0u);
    lo = (byte + 8 < z) ? base[byte + 8] : 0u;
    const uint32_t sh = (uint32_t)(bit & 7u);
    return sh ? ((hi << sh) | (lo >> (8 - sh))) : hi;
}

__global__ void __launch_bounds__(128)
k_dec_stitch(const uint8_t *__restrict__ payload, DecPlan plan, DecTablesGlobal g, uint16_t *__restrict__ sub_info,
             uint32_t *__restrict__ seq_end, uint32_t *__restrict__ seq_count, uint64_t nseq,
             uint32_t *__restrict__ changed_flag, const uint32_t *__restrict__ prev_changed)
{
    // a pass in which no sequence end moved leaves every sequence consistent with its predecessor: the pass after it
    // (launched blind, without a host round trip) has nothing to do
    if (prev_changed && *prev_changed == 0) return;
    const uint64_t seq = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (seq >= nseq) return;
    const uint32_t b = find_block(plan.seq0, plan.nb, seq);
    const uint64_t sib = seq - plan.seq0[b];
    if (sib == 0) return;  // first sequence of a block starts at bit 0 by definition
    uint32_t cur = seq_end[seq - 1];
    const uint64_t z = plan.z[b];
    const uint8_t *base = payload + plan.pay_off[b];
    DecTables t;
    t.lut = g.lut + (uint64_t)b * DEC_LUT_SIZE;
    t.sleft = g.sleft + (uint64_t)b * 256;
    t.slen = g.slen + (uint64_t)b * 256;
    t.ssym = g.ssym + (uint64_t)b * 256;
    const uint64_t seq_bit0 = sib * DEC_SEQ_BYTES * 8;
    const uint64_t total_bits = z * 8;
    bool end_moved = false;
    int32_t delta = 0;  // change of this sequence's symbol count
    for (uint32_t k = 0; k < DEC_THREADS; k++) {
        const uint64_t gi = seq * DEC_THREADS + k;
        const uint16_t rec = sub_info[gi];
        if (sub_start(rec) == cur) break;  // met the recorded state: everything after is already right
        const uint64_t lo = seq_bit0 + (uint64_t)k * DEC_SUB_BITS;
        const uint64_t hi = lo + DEC_SUB_BITS;
        const uint64_t stop = hi < total_bits ? hi : total_bits;
        uint64_t q = lo + cur;
        uint32_t count = 0;
        while (q < stop) {
            uint32_t sym;
            uint32_t l = dec_symbol(t, window_global(base, z, q), &sym);
            if (!l) l = 1;
            q += l;
            count++;
        }
        sub_info[gi] = sub_pack(cur, count);
        delta += (int32_t)count - (int32_t)sub_count(rec);
        cur = q > hi ? (uint32_t)(q - hi) : 0u;
        if (k == DEC_THREADS - 1 && seq_end[seq] != cur) {
            seq_end[seq] = cur;
            end_moved = true;
        }
    }
    if (delta) seq_count[seq] = (uint32_t)((int32_t)seq_count[seq] + delta);
    // Only a moved sequence end can invalidate what another thread of this pass decided: a pass in which
    // no end moved leaves every sequence consistent with its predecessor, and nothing is left to do.
    if (end_moved) atomicOr(changed_flag, 1u);
}

// ---- phase 3: output offsets ---------------------------------------------------------------------
// one CTA per block: exclusive scan of its sequences' counts; checks the block holds >= n symbols
__global__ void __launch_bounds__(256)
k_dec_scan(DecPlan plan, const uint32_t *__restrict__ seq_count, uint64_t *__restrict__ seq_excl,
           uint32_t *__restrict__ err, const uint32_t *__restrict__ still_moving)
{
    // the stitch passes ran without the host looking at their flag: if the last one still moved a sequence end, the
    // counts below are not final -- say so, the caller decodes again with the polling loop
    if (blockIdx.x == 0 && threadIdx.x == 0 && still_moving && *still_moving) atomicOr(err, (uint32_t)ERR_STITCH);
    __shared__ uint32_t scan[9];
    __shared__ unsigned long long s_carry;
    const uint32_t b = blockIdx.x;
    const uint64_t s0 = plan.seq0[b], s1 = plan.seq0[b + 1];
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    for (uint64_t base = s0; base < s1; base += 256) {
        const uint64_t i = base + threadIdx.x;
        const uint32_t v = i < s1 ? seq_count[i] : 0u;
        uint32_t total;
        const uint32_t excl = block_excl_scan<256>(v, scan, &total);
        if (i < s1) seq_excl[i] = s_carry + excl;
        __syncthreads();
        if (threadIdx.x == 0) s_carry += total;
        __syncthreads();
    }
    if (threadIdx.x == 0 && s_carry < plan.n[b]) atomicOr(err, 8u);  // payload too short for n symbols
}

// ---- phase 4: final decode + write --------------------------------------------------------------
// Symbols of a sequence are staged in shared memory (co-aligned with their destination) and
// copied out with 16-byte stores; a sequence that decodes to more than DEC_STAGE symbols
// (codes of 1-2 bits) writes straight to global memory instead.
struct DecWriteShared {  // no lut1 (lut3 entries carry the first code's length, core.h dec_lut3_entry) and no payload rows (a thread only
                         // ever decodes its own sub-sequence here: six words in registers): 31 KB
    uint64_t sleft[256];
    uint8_t slen[256];
    uint8_t ssym[256];
    uint32_t scan[DEC_THREADS / 32 + 1];
    alignas(16) uint32_t lut3[DEC_LUT_SIZE];
    alignas(16) uint8_t stage[DEC_STAGE + 48];  // + the destination's misalignment (< 16)
};

// Sink of the write pass: the thread's symbols go to consecutive shared byte addresses from `base`.  CHECKED
// drops the symbols at or past `room` (padding bits after a block's last symbol decode to garbage).
template <bool CHECKED>
struct StageSink {
    uint32_t base, room;
    uint32_t addr;  // where the next symbol goes
    __device__ __forceinline__ void put_if(uint32_t a, uint32_t sym, uint32_t p) const  // store if p != 0
    {
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.u8 [%0], %1;\n\t}" ::"r"(a), "r"(sym), "r"(p) : "memory");
    }
    __device__ __forceinline__ void one(uint32_t sym)
    {
        if (CHECKED)
            put_if(addr, sym, addr - base < room);
        else
            asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(sym) : "memory");
        addr++;
    }
    __device__ __forceinline__ void two(uint32_t e3)
    {
        const uint32_t second = e3 & 0x20000u;
        if (CHECKED) {
            put_if(addr, e3, addr - base < room);
            put_if(addr + 1, e3 >> 8, (addr + 1 - base < room) ? second : 0u);
        } else {
            asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(e3) : "memory");
            put_if(addr + 1, e3 >> 8, second);
        }
        addr += (e3 >> 16) & 3u;  // one or two symbols
    }
    __device__ __forceinline__ uint32_t count() const { return addr - base; }
};

#ifndef SHAFA_DECW_MIN_CTAS
#define SHAFA_DECW_MIN_CTAS 5
#endif
__global__ void __launch_bounds__(DEC_THREADS, SHAFA_DECW_MIN_CTAS)
k_dec_write(const uint8_t *__restrict__ payload, DecPlan plan, DecTablesGlobal g,
            const uint16_t *__restrict__ sub_info, const uint64_t *__restrict__ seq_excl, uint8_t *__restrict__ out,
            uint64_t nseq, uint32_t *__restrict__ err)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DecWriteShared &ws = *reinterpret_cast<DecWriteShared *>(smem_raw);
    DecWriteShared &sh = ws;
    const unsigned tid = threadIdx.x;
    uint64_t seq_first, seq_last;
    dec_chunk(nseq, &seq_first, &seq_last);
    uint32_t cur_block = 0xffffffffu;
    uint32_t bad = 0;
    if (seq_first >= seq_last) return;
    SeqCursor cursor;
    cursor.init(plan.seq0, plan.nb, seq_first);
    SeqFetch cur = dec_fetch_seq(payload, plan, cursor, seq_first);
    uint16_t rec = sub_info[seq_first * DEC_THREADS + tid];
    uint64_t sfirst = seq_excl[seq_first];  // index of the sequence's first symbol inside the block
    for (uint64_t seq = seq_first; seq < seq_last; seq++) {
        // the next sequence's global loads are in flight while this one is decoded
        SeqFetch nxt = cur;
        uint16_t rec_nxt = 0;
        uint64_t sfirst_nxt = 0;
        if (seq + 1 < seq_last) {
            nxt = dec_fetch_seq(payload, plan, cursor, seq + 1);
            rec_nxt = sub_info[(seq + 1) * DEC_THREADS + tid];
            sfirst_nxt = seq_excl[seq + 1];
        }
        const uint32_t b = cur.b;
        const uint64_t sib = cur.sib, z = cur.z;
        if (b != cur_block) {
            cur_block = b;
            __syncthreads();  // nobody decodes with the old tables any more
            sh.sleft[tid] = g.sleft[(uint64_t)b * 256 + tid];
            sh.slen[tid] = g.slen[(uint64_t)b * 256 + tid];
            sh.ssym[tid] = g.ssym[(uint64_t)b * 256 + tid];
            const uint4 *src3 = reinterpret_cast<const uint4 *>(g.lut3 + (uint64_t)b * DEC_LUT_SIZE);
            uint4 *dst3 = reinterpret_cast<uint4 *>(ws.lut3);
            for (int i = tid; i < DEC_LUT_SIZE * 4 / 16; i += DEC_THREADS) dst3[i] = src3[i];
        }
        uint32_t words[DEC_ROW_WORDS];
        dec_row_words(cur.row, z, sib, words);
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");  // the previous sequence's bulk store has read the stage
        uint32_t total;
        // the scan's barriers also fence the table loads above and the previous sequence's copy-out
        const uint32_t excl = block_excl_scan<DEC_THREADS>(sub_count(rec), sh.scan, &total);
        DecViewS t;
        t.pairs_s = 0u;
        t.lut1_s = 0u;
        t.lut2_s = 0u;
        t.lut3_s = smem_u32(ws.lut3);
        asm volatile("" : "+r"(t.lut3_s));  // not rebuilt from the shared window inside the loops
        t.lut1 = nullptr;
        t.sleft = sh.sleft;
        t.slen = sh.slen;
        t.ssym = sh.ssym;
        const uint64_t bits_left = (z - sib * DEC_SEQ_BYTES) * 8;
        const uint32_t limit = bits_left < (uint64_t)DEC_SEQ_BYTES * 8 ? (uint32_t)bits_left : DEC_SEQ_BYTES * 8;
        const uint32_t sub_lo = tid * DEC_SUB_BITS;
        const uint32_t stop = limit > sub_lo ? (limit - sub_lo < (uint32_t)DEC_SUB_BITS ? limit - sub_lo : (uint32_t)DEC_SUB_BITS) : 0u;
        const uint64_t n = plan.n[b];
        if (sfirst < n) {  // else: only padding garbage from here on (uniform for the CTA)
            const uint64_t seq_room = n - sfirst;  // symbols of this sequence that are real
            const uint32_t seq_syms = seq_room < total ? (uint32_t)seq_room : total;
            uint8_t *gdst = out + plan.out_off[b] + sfirst;
            const uint32_t room = excl < seq_syms ? seq_syms - excl : 0u;
            if (seq_syms <= (uint32_t)DEC_STAGE) {
                // the sequence's place in the output is known before it is decoded (k_dec_scan), so the symbols are staged
                // CO-ALIGNED with their destination: stage byte (gdst & 15) + i is output byte i, and everything between
                // the first and the last 16-byte boundary leaves through ONE bulk async store (cp.async.bulk shared ->
                // global, "TMA" 1-D) issued by one thread
                const uint32_t mis = (uint32_t)((uintptr_t)gdst & 15u);
                uint32_t dst_s = smem_u32(ws.stage) + mis + excl;  // explicit st.shared: no generic address in the loop
                asm volatile("" : "+r"(dst_s));
                DecSub r;
                if (seq_room >= total) {  // every symbol of this sequence is real (all but a block's last sequences)
                    StageSink<false> sink{dst_s, room, dst_s};
                    r = dec_emit_words(t, words, sub_start(rec), stop, sink);
                } else {
                    StageSink<true> sink{dst_s, room, dst_s};
                    r = dec_emit_words(t, words, sub_start(rec), stop, sink);
                }
                bad |= r.bad && room;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // my stores to the stage, before the bulk copy reads them
                __syncthreads();
                {
                    const uint32_t head = mis ? (16u - mis) : 0u;                       // bytes before the first 16-byte boundary
                    const uint32_t h = head < seq_syms ? head : seq_syms;
                    const uint32_t body = (seq_syms - h) & ~15u;                         // whole 16-byte groups
                    if (tid == 0 && body)
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;\n\tcp.async.bulk.commit_group;" ::"l"(gdst + h),
                                     "r"(smem_u32(ws.stage) + mis + h), "r"(body)
                                     : "memory");
                    if (tid >= 32 && tid < 32 + h) gdst[tid - 32] = ws.stage[mis + tid - 32];
                    const uint32_t done = h + body;
                    if (tid >= 64 && tid < 64 + (seq_syms - done)) gdst[done + tid - 64] = ws.stage[mis + done + tid - 64];
                }
            } else {
                uint8_t *dst = gdst + excl;
                auto put = [&](uint32_t i, uint32_t sym) {
                    if (i < room) dst[i] = (uint8_t)sym;
                };
                DecFnSink<decltype(put)> sink{put, 0u};
                const DecSub r = dec_emit_words(t, words, sub_start(rec), stop, sink);
                bad |= r.bad && room;
            }
        }
        cur = nxt;
        rec = rec_nxt;
        sfirst = sfirst_nxt;
    }
    if (tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");  // the last bulk store has completed
    // a bad window only matters if it produced one of the n real symbols
    if (bad) atomicOr(err, (uint32_t)ERR_BAD_WINDOW);
}

// ---- launcher -------------------------------------------------------------------------------------
// plan arrays of one path: pay_off[nb] z[nb] n[nb] out_off[nb] seq0[nb+1]; blocks of the OTHER path get z = n = 0
// (no sequences), so one output buffer is filled by both.
static uint64_t fill_plan(uint64_t *hp, const uint64_t *h_payload_off, const uint64_t *h_z, const uint64_t *h_n, uint64_t nb,
                          const uint8_t *flags, bool want_fused)
{
    uint64_t nseq = 0, out_off = 0;
    for (uint64_t b = 0; b < nb; b++) {
        const bool mine = flags ? (dec_block_fused_ok(flags[b]) == want_fused) : !want_fused;
        hp[b] = h_payload_off[b];
        hp[nb + b] = mine ? h_z[b] : 0;
        hp[2 * nb + b] = mine ? h_n[b] : 0;
        hp[3 * nb + b] = out_off;
        hp[4 * nb + b] = nseq;
        out_off += h_n[b];
        if (mine) nseq += (h_z[b] + DEC_SEQ_BYTES - 1) / DEC_SEQ_BYTES;
    }
    hp[5 * nb] = nseq;
    return nseq;
}
static DecPlan plan_at(const uint64_t *dp, uint64_t nb)
{
    DecPlan plan;
    plan.pay_off = dp;
    plan.z = dp + nb;
    plan.n = dp + 2 * nb;
    plan.out_off = dp + 3 * nb;
    plan.seq0 = dp + 4 * nb;
    plan.nb = (uint32_t)nb;
    plan.cls_seq0[0] = plan.cls_seq0[1] = nullptr;
    plan.lean_slot = nullptr;
    return plan;
}

// the exact path: sync / stitch / scan / write over the blocks of `plan`
// n_big: how many of the plan's blocks have enough payload to be lean (dec_block_lean; whether their code is of that
// kind is only known on the device): the number of 64 KiB table slots to provide
static int launch_sf_decode_exact(shafa_cuda_ctx *ctx, const uint8_t *d_payload, const DecPlan &plan_in, const DecTablesGlobal &g_in,
                                  uint64_t nb, uint64_t nseq, uint64_t n_big, uint8_t *d_sym, uint32_t *d_err, bool careful_stitch)
{
    if (nseq == 0) return SHAFA_SUCCESS;
    DecPlan plan = plan_in;
    DecTablesGlobal g = g_in;
    {
        // the two sequence classes (lean blocks / everything else) and the lean blocks' tables, all decided on the device
        uint64_t *cls = (uint64_t *)scratch_get(ctx, SCR_DEC_CLS, 2 * (nb + 1) * sizeof(uint64_t) + nb * sizeof(uint32_t));
        if (!cls) return SHAFA_LACK_OF_MEMORY;
        plan.cls_seq0[0] = cls;
        plan.cls_seq0[1] = cls + nb + 1;
        uint32_t *slot = (uint32_t *)(cls + 2 * (nb + 1));
        plan.lean_slot = slot;
        SHAFA_LAUNCH(ctx, "dec_classes", k_dec_classes, dim3(1), dim3(256), 0, plan.seq0, plan.z, (const uint8_t *)g.flags, (uint32_t)nb, cls,
                     cls + nb + 1, slot);
        g.lut16 = nullptr;
        if (n_big) {
            g.lut16 = (uint8_t *)scratch_get(ctx, SCR_DEC_LUT16, n_big * (size_t)DEC_LUT16_SIZE);
            if (!g.lut16) return SHAFA_LACK_OF_MEMORY;
            SHAFA_LAUNCH(ctx, "dec_lut16", k_dec_lut16, dim3((unsigned)nb), dim3(256), 0, plan, g);
        }
        SHAFA_CHECK_LAUNCH(ctx);
    }
    uint16_t *sub_info = (uint16_t *)scratch_get(ctx, SCR_DEC_SUB, nseq * DEC_THREADS * sizeof(uint16_t));
    // seq arrays: seq_end u32[nseq], seq_count u32[nseq], seq_excl u64[nseq]
    uint8_t *seqbuf = (uint8_t *)scratch_get(ctx, SCR_DEC_SEQ, nseq * 16 + 64);
    uint32_t *flag = misc_ptr(ctx, MISC_FLAG);
    if (!sub_info || !seqbuf || !flag) return SHAFA_LACK_OF_MEMORY;
    uint64_t *seq_excl = (uint64_t *)seqbuf;
    uint32_t *seq_end = (uint32_t *)(seqbuf + nseq * 8);
    uint32_t *seq_count = seq_end + nseq;

    int &occ_sync = ctx->occ[OCC_DEC_SYNC], &occ_write = ctx->occ[OCC_DEC_WRITE];
    const size_t smem_w = sizeof(DecWriteShared);
    if (!occ_sync) {
        SHAFA_CUDA_TRY(ctx, cudaFuncSetAttribute(k_dec_write, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_w));
        SHAFA_CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_sync, k_dec_sync, DEC_THREADS, 0));
        SHAFA_CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_write, k_dec_write, DEC_THREADS, smem_w));
        if (occ_sync < 1) occ_sync = 1;
        if (occ_write < 1) occ_write = 1;
    }
    // persistent CTAs, each owning a contiguous chunk of sequences (4 chunks per resident CTA for balance)
    uint64_t grid_sync = (uint64_t)ctx->sm_count * occ_sync * 4, grid_write = (uint64_t)ctx->sm_count * occ_write * 4;
    if (grid_sync > nseq) grid_sync = nseq;
    if (grid_write > nseq) grid_write = nseq;
    SHAFA_LAUNCH(ctx, "dec_sync", k_dec_sync, dim3((unsigned)grid_sync), dim3(DEC_THREADS), 0, d_payload, plan, g, sub_info,
                 seq_end, seq_count);
    if (n_big) {
        int &occ_lean = ctx->occ[OCC_DEC_LEAN];
        const size_t smem_l = sizeof(DecLeanShared);
        if (!occ_lean) {
            SHAFA_CUDA_TRY(ctx, cudaFuncSetAttribute(k_dec_sync_lean, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_l));
            SHAFA_CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ_lean, k_dec_sync_lean, LEAN_THREADS, smem_l));
            if (occ_lean < 1) occ_lean = 1;
        }
        uint64_t grid_lean = (uint64_t)ctx->sm_count * occ_lean * 4;
        if (grid_lean > nseq) grid_lean = nseq;
        SHAFA_LAUNCH(ctx, "dec_sync_lean", k_dec_sync_lean, dim3((unsigned)grid_lean), dim3(LEAN_THREADS), smem_l, d_payload, plan, g,
                     sub_info, seq_end, seq_count);
    }
    SHAFA_CHECK_LAUNCH(ctx);
    // Stitch across chunk boundaries until a pass moves no sequence end.  The first pass almost always settles it and the
    // second then moves nothing, so the normal path launches two passes back to back with NO host round trip in
    // between; k_dec_scan raises ERR_STITCH if the second pass still moved an end, and only then (careful_stitch) the
    // passes are repeated under host polling.
    uint32_t *h_flag = (uint32_t *)pinned_get(ctx, 64);
    if (!h_flag) return SHAFA_LACK_OF_MEMORY;
    const uint32_t *still_moving = nullptr;
    if (!careful_stitch) {
        SHAFA_CUDA_TRY(ctx, cudaMemsetAsync(flag, 0, 8, ctx->stream));
        for (int pass = 0; pass < 2; pass++)
            SHAFA_LAUNCH(ctx, "dec_stitch", k_dec_stitch, dim3((unsigned)((nseq + 127) / 128)), dim3(128), 0, d_payload,
                         plan, g, sub_info, seq_end, seq_count, nseq, flag + pass, pass ? (const uint32_t *)flag : (const uint32_t *)nullptr);
        SHAFA_CHECK_LAUNCH(ctx);
        still_moving = flag + 1;
    }
    for (uint64_t pass = 0; careful_stitch; pass++) {
        SHAFA_CUDA_TRY(ctx, cudaMemsetAsync(flag, 0, 4, ctx->stream));
        SHAFA_LAUNCH(ctx, "dec_stitch", k_dec_stitch, dim3((unsigned)((nseq + 127) / 128)), dim3(128), 0, d_payload,
                     plan, g, sub_info, seq_end, seq_count, nseq, flag, (const uint32_t *)nullptr);
        SHAFA_CHECK_LAUNCH(ctx);
        SHAFA_CUDA_TRY(ctx, cudaMemcpyAsync(h_flag, flag, 4, cudaMemcpyDeviceToHost, ctx->stream));
        SHAFA_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        if (*h_flag == 0) break;
        if (pass > nseq + 2) return set_error(ctx, SHAFA_FILE_UNRECOGNIZABLE, "sf_decode: stitching did not converge");
    }
    SHAFA_LAUNCH(ctx, "dec_scan", k_dec_scan, dim3((unsigned)nb), dim3(256), 0, plan, seq_count, seq_excl, d_err, still_moving);
    SHAFA_LAUNCH(ctx, "dec_write", k_dec_write, dim3((unsigned)grid_write), dim3(DEC_THREADS), smem_w, d_payload, plan, g,
                 sub_info, seq_excl, d_sym, nseq, d_err);
    SHAFA_CHECK_LAUNCH(ctx);
    return SHAFA_SUCCESS;
}

// force_exact: every block through the exact path (set after the fused kernel reported ERR_SPECULATION, or by
// SHAFA_DECODE_EXACT=1 for A/B measurements)
int launch_sf_decode(shafa_cuda_ctx *ctx, const uint8_t *d_payload, const uint64_t *h_payload_off,
                     const uint64_t *h_z, const uint64_t *h_n, uint64_t nb, const uint8_t *d_len,
                     const uint64_t *d_code, uint8_t *d_sym, uint32_t *d_err, bool force_exact, bool careful_stitch)
{
    if (nb == 0) return SHAFA_SUCCESS;
    {
        // The fused single-pass kernel (sf_decode_fused.cu) is correct and tested, but on B200 it is latency-bound at
        // three resident CTAs per SM (72 KB of records per CTA) and 2.4x slower than this file's multi-kernel path
        // (profiles/r02_fused_decoder.md): it runs only when asked for.
        const char *e = getenv("SHAFA_DECODE_FUSED");
        if (!(e && *e && *e != '0')) force_exact = true;
    }
    uint64_t any = 0;
    for (uint64_t b = 0; b < nb; b++) {
        if (h_payload_off[b] & 15) return set_error(ctx, SHAFA_FILE_STREAM_FAILED, "sf_decode: payload offset not 16-byte aligned");
        // A one-symbol block is written as "@0@" by module C and its .cod block holds 256 empty strings: which
        // symbol it was is not in the files (the reference decoder dereferences NULL there, d.c:533-536).
        if (h_n[b] != 0 && h_z[b] == 0) return set_error(ctx, SHAFA_FILE_UNRECOGNIZABLE, "sf_decode: block %llu has %llu symbols but an empty payload (one-symbol block: not decodable from .cod/.shaf)", (unsigned long long)b, (unsigned long long)h_n[b]);
        any += (h_z[b] + DEC_SEQ_BYTES - 1) / DEC_SEQ_BYTES;
    }
    if (any == 0) return SHAFA_SUCCESS;
    if (any > 0x7fffffffull) return set_error(ctx, SHAFA_LACK_OF_MEMORY, "sf_decode: too many sequences");

    // tables
    const size_t lut_bytes = nb * DEC_LUT_SIZE * sizeof(uint16_t);
    const size_t tab_bytes = 6 * lut_bytes + nb * 256 * (sizeof(uint64_t) + 2) + nb + 64;
    uint8_t *tab = (uint8_t *)scratch_get(ctx, SCR_DEC_LUT, tab_bytes);
    if (!tab) return SHAFA_LACK_OF_MEMORY;
    DecTablesGlobal g;
    g.sleft = (uint64_t *)tab;
    uint8_t *p = tab + nb * 256 * sizeof(uint64_t);
    g.lut = (uint16_t *)p;   p += lut_bytes;
    g.lut2 = p;              p += lut_bytes;  // DEC_LUT2_SIZE bytes per block = lut_bytes per block
    g.lut3 = (uint32_t *)p;  p += 2 * lut_bytes;
    g.lutp = (uint32_t *)p;  p += 2 * lut_bytes;
    g.slen = p;              p += nb * 256;
    g.ssym = p;              p += nb * 256;
    g.flags = p;
    SHAFA_LAUNCH(ctx, "dec_tables", k_dec_tables, dim3((unsigned)nb), dim3(256), 0, d_len, d_code, g, force_exact ? 1 : 0);
    SHAFA_CHECK_LAUNCH(ctx);

    // which path each block takes is a property of its code table, decided on the device; the plans are built here
    const uint8_t *h_flags = nullptr;
    if (!force_exact) {
        uint8_t *hf = (uint8_t *)pinned_get(ctx, nb);
        if (!hf) return SHAFA_LACK_OF_MEMORY;
        SHAFA_CUDA_TRY(ctx, cudaMemcpyAsync(hf, g.flags, nb, cudaMemcpyDeviceToHost, ctx->stream));
        SHAFA_CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        h_flags = hf;
    }
    const size_t plan_words = 5 * nb + 1;
    uint64_t *hp = (uint64_t *)pinned_get(ctx, 2 * plan_words * sizeof(uint64_t));
    uint64_t *dp = (uint64_t *)scratch_get(ctx, SCR_PLAN, 2 * plan_words * sizeof(uint64_t));
    if (!hp || !dp) return SHAFA_LACK_OF_MEMORY;
    const uint64_t nseq_fused = h_flags ? fill_plan(hp, h_payload_off, h_z, h_n, nb, h_flags, true) : 0;
    const uint64_t nseq_exact = fill_plan(hp + plan_words, h_payload_off, h_z, h_n, nb, h_flags, false);
    SHAFA_CUDA_TRY(ctx, cudaMemcpyAsync(dp, hp, 2 * plan_words * sizeof(uint64_t), cudaMemcpyHostToDevice, ctx->stream));
    int rc = SHAFA_SUCCESS;
    if (nseq_fused) {
        const DecPlan plan = plan_at(dp, nb);
        DecSeqRec *rec = (DecSeqRec *)scratch_get(ctx, SCR_DEC_REC, nseq_fused * sizeof(DecSeqRec));
        if (!rec) return SHAFA_LACK_OF_MEMORY;
        SHAFA_LAUNCH(ctx, "dec_seqrec", k_dec_seqrec, dim3((unsigned)nb), dim3(256), 0, plan, rec);
        SHAFA_CHECK_LAUNCH(ctx);
        if ((rc = launch_sf_decode_fused(ctx, d_payload, plan, g, rec, nseq_fused, d_sym, d_err))) return rc;
    }
    if (nseq_exact) {
        uint64_t n_big = 0;  // blocks of the exact plan that are large enough to be lean (fill_plan zeroes z of the other path's blocks)
        for (uint64_t b = 0; b < nb; b++) n_big += hp[plan_words + nb + b] >= DEC_LEAN_MIN_Z ? 1 : 0;
        rc = launch_sf_decode_exact(ctx, d_payload, plan_at(dp + plan_words, nb), g, nb, nseq_exact, n_big, d_sym, d_err, careful_stitch);
    }
    return rc;
}

}  // namespace shafa
