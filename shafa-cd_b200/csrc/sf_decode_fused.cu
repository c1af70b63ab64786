// sf_decode_fused.cu -- Shannon-Fano prefix decode (reference create_tree +
// shafa_block_decompressor, d.c:422-551) as ONE single-pass, warp-specialised kernel.
//
// The .shaf bitstream has no sync points, so a block is cut into 4 KiB "sequences" of 256 sub-sequences
// of 128 bits (one thread each) and decoded speculatively (self-synchronising decoding, Weissenberger &
// Schmidt ICPP'18) -- but every payload bit is decoded about 1.2 times, not once per round plus once more
// for the output (core.h, "Fused single-pass decoder"):
//
//   * persistent CTAs take sequences in stream order through a ticket counter, so a sequence's
//     predecessors are always running or done (no wait on a CTA that is not resident);
//   * 8 DECODE warps + 1 CHAIN warp per CTA.  Everything with global-memory latency in its dependency
//     chain belongs to the chain warp and runs beside the decoding instead of in front of it: taking the
//     next ticket, the bulk async copy (cp.async.bulk + mbarrier, "TMA" 1-D) of the NEXT sequence's
//     payload into shared memory, polling the predecessor sequence's end offset, and the decoupled
//     look-back (common.cuh) over the symbol counts.  The two sides talk through mbarriers;
//   * pass 1: every thread decodes its sub-sequence from bit 0 with a two-codes-per-lookup table and
//     keeps the symbols as records in shared memory;
//   * rounds: a thread whose true start (its predecessor's end) differs walks single codes from there
//     until it meets its own record list, and keeps head + list tail.  Ends that change ripple on
//     (rare); the loop ends when every start equals the predecessor's end;
//   * the sequence's end offset is published as soon as that holds with the first sub-sequence started
//     at bit 0: a prefix code re-synchronises within a few codes, so the end of a 4 KiB sequence does
//     not depend on where its first code started.  The true start then comes from the predecessor's
//     descriptor, the first sub-sequence is corrected, and the published end is VERIFIED -- if it ever
//     moved (codes that do not self-synchronise), the kernel raises ERR_SPECULATION and the caller
//     decodes again with the exact multi-kernel path (sf_decode.cu), so the result never depends on the guess;
//   * the symbols are compacted into a staging buffer; where they go (the look-back result) is only
//     needed one sequence later: the copy-out of sequence k runs in the middle of sequence k + 1,
//     between its speculative rounds and the arrival of its true start, where it hides both waits.
//
// Blocks whose code table is not a complete prefix-free code, or looks like it does not self-synchronise
// (nearly one code length), never come here: k_dec_tables flags them and launch_sf_decode sends them
// through the exact path.  Algorithmic bytes: Z (payload in) + S (symbols out).
#include <stdlib.h>

#include "sf_decode.cuh"

namespace shafa {

constexpr int DECF_STAGE = 10240;               // symbols of one sequence staged before the coalesced copy-out
constexpr uint32_t DECF_RAW_BYTES = DEC_SEQ_BYTES + 16;
constexpr int DECF_CTA = DEC_THREADS + 32;      // 8 decode warps + the chain warp
constexpr uint32_t DECF_SPIN_LIMIT = 1u << 22;  // polls before a wait gives up (ERR_TIMEOUT): no wait in here can hang the GPU

struct DecfInfo {     // a ticket: which sequence, and where its payload lies
    uint64_t src_off;
    uint32_t seq, b, bytes, pad;
};

struct DecfShared {
    uint64_t pairs[DEC_PAIR_SLOTS];                   // row of sub-sequence s at [s * DEC_ROW, (s + 1) * DEC_ROW)
    alignas(16) uint8_t stage[DECF_STAGE + 32];       // + the read-ahead of copy_out_unaligned
    alignas(16) uint32_t lutp[DEC_LUT_SIZE];          // pair LUT of the current block
    alignas(16) uint64_t sleft[256];
    alignas(16) uint8_t slen[256];
    alignas(16) uint8_t ssym[256];
    uint32_t slots[DEC_THREADS * DECF_CAP];           // record j of thread t at [j * 256 + t]: one bank per lane
    uint8_t heads[DEC_THREADS * DECF_HEAD];           // head symbol i of thread t at [i * 256 + t]
    alignas(16) uint8_t raw[DECF_RAW_BYTES];          // landing zone of the bulk copy
    uint8_t ends[DEC_THREADS];                        // overshoot of sub-sequence i into sub-sequence i + 1
    uint32_t scan[DEC_THREADS / 32 + 1];
    alignas(8) uint64_t mb_payload;                   // chain -> decode: payload of the next sequence has landed (tx bytes)
    uint64_t mb_rawfree;                              // decode -> chain: rows built, the landing zone may be overwritten
    uint64_t mb_carry;                                // chain -> decode: sh.carry holds the predecessor sequence's end
    uint64_t mb_total;                                // decode -> chain: sh.total holds this sequence's symbol count
    uint64_t mb_prefix;                               // chain -> decode: sh.prefix[k & 1] holds the look-back result
    uint64_t prefix[2];
    DecfInfo info[2];
    uint32_t total;
    uint32_t carry;
    uint32_t carry_early;                             // decode thread 0 -> decode warps: the true start is in use from round 1
    uint32_t dead;                                    // a wait timed out: every later wait falls through
};

// ---- mbarrier + bulk async copy (TMA 1-D) ------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void bulk_load(void *smem_dst, const void *gsrc, uint32_t bytes, uint64_t *bar)
{
    const uint32_t b = smem_u32(bar);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(bytes) : "memory");
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
                 "l"(gsrc), "r"(bytes), "r"(b)
                 : "memory");
}
__device__ __forceinline__ bool mbar_test(uint64_t *bar, uint32_t parity)
{
    uint32_t done;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(done)
                 : "r"(smem_u32(bar)), "r"(parity)
                 : "memory");
    return done != 0;
}
// bounded wait: after DECF_SPIN_LIMIT polls the kernel is declared dead (ERR_TIMEOUT) and every wait falls through
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity, volatile uint32_t *dead, uint32_t *err)
{
    uint32_t spins = 0;
    while (!mbar_test(bar, parity)) {
        if (*dead) return;
        if (++spins > DECF_SPIN_LIMIT) {
            *dead = 1;
            atomicOr(err, (uint32_t)ERR_TIMEOUT);
            return;
        }
    }
}
// barriers of the 256 decode threads only (the chain warp never joins them)
__device__ __forceinline__ void dbar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }
__device__ __forceinline__ bool dbar_or(bool pred)
{
    uint32_t r;
    asm volatile("{\n\t.reg .pred p, q;\n\tsetp.ne.u32 p, %1, 0;\n\tbar.red.or.pred q, 1, 256, p;\n\tselp.u32 %0, 1, 0, q;\n\t}"
                 : "=r"(r)
                 : "r"((uint32_t)pred)
                 : "memory");
    return r != 0;
}
__device__ __forceinline__ uint32_t decode_excl_scan(uint32_t v, uint32_t *smem, uint32_t *total)
{
    constexpr int NW = DEC_THREADS / 32;
    const uint32_t incl = warp_incl_scan(v);
    if (lane_id() == 31) smem[warp_id()] = incl;
    dbar();
    if (warp_id() == 0) {
        uint32_t w = (lane_id() < NW) ? smem[lane_id()] : 0u;
        uint32_t wi = warp_incl_scan(w);
        if (lane_id() < NW) smem[lane_id()] = wi - w;
        if (lane_id() == NW - 1) smem[NW] = wi;
    }
    dbar();
    *total = smem[NW];
    return smem[warp_id()] + incl - v;
}

// view of the tables / slots for core.h's decf_* functions: everything through 32-bit shared addresses
struct DecfViewS {
    uint32_t pairs_s, lutp_s;
    const uint64_t *sleft;
    const uint8_t *slen;
    const uint8_t *ssym;
    __device__ __forceinline__ uint64_t pair(uint32_t slot) const
    {
        uint64_t v;
        asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(pairs_s + slot * 8u));
        return v;
    }
    __device__ __forceinline__ uint32_t lp(uint32_t i) const
    {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(lutp_s + i * 4u));
        return v;
    }
    __device__ __forceinline__ uint32_t ld32s(uint32_t a) const
    {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory");
        return v;
    }
    __device__ __forceinline__ void st32s(uint32_t a, uint32_t v) const { asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
};
// core.h addresses records as slot + 4 j and head bytes as head + i; the interleaved layout scales both by 256
struct DecfViewI : DecfViewS {
    uint32_t slot0, head0;  // this thread's record 0 / head byte 0
    __device__ __forceinline__ uint32_t ld32(uint32_t a) const { return ld32s(slot0 + (a - slot0) * 256u); }
    __device__ __forceinline__ void st32(uint32_t a, uint32_t v) const { st32s(slot0 + (a - slot0) * 256u, v); }
    __device__ __forceinline__ void st8(uint32_t a, uint32_t v) const { sts_u8(head0 + (a - head0) * 256u, v); }
};

struct DecfSub {
    DecfMain mn;
    uint32_t start, end, count, hn, jsp;
};

__device__ __forceinline__ uint32_t decf_over(uint32_t end) { return end > (uint32_t)DEC_SUB_BITS ? end - DEC_SUB_BITS : 0u; }

__device__ __forceinline__ void decf_restart(const DecfViewI &t, uint32_t row, uint32_t s, uint32_t stop, DecfSub &u)
{
    u.start = s;
    const DecfSplice sp = decf_splice(t, row, t.slot0, t.head0, u.mn, s, stop);
    const uint32_t mrec = u.mn.norec ? 0u : u.mn.m;
    if (sp.kind == 0) {  // met the list: head + tail of the list
        u.hn = sp.hn;
        u.jsp = sp.jsp;
        u.count = sp.hn + (u.mn.k - sp.kskip);
        u.end = u.mn.end;
    } else if (sp.kind == 1) {  // left the sub-sequence without meeting it: the head is everything
        u.hn = sp.hn;
        u.jsp = mrec;
        u.count = sp.hn;
        u.end = sp.end;
    } else {  // no meeting point within DECF_HEAD codes: the list is rebuilt from the new start
        u.mn = decf_pass(t, row, t.slot0, s, stop);
        u.hn = 0;
        u.jsp = 0;
        u.count = u.mn.k;
        u.end = u.mn.end;
    }
}

// look-back over the symbol counts with a 256-descriptor window (8 per lane): the chain warp has one sequence's
// decoding time for it, and the nearest published prefix is a few hundred sequences back
__device__ __forceinline__ uint64_t decf_lookback(const uint64_t *desc, int64_t tile, int64_t first, volatile uint32_t *dead, uint32_t *err)
{
    constexpr int PER_LANE = 8;
    const unsigned lane = lane_id();
    uint64_t excl = 0;
    int64_t pos = tile - 1;
    uint32_t spins = 0;
    for (;;) {
        uint64_t d[PER_LANE];
        bool pending;
        do {
            pending = false;
#pragma unroll
            for (int j = 0; j < PER_LANE; j++) {
                const int64_t idx = pos - (int64_t)lane - 32 * j;
                d[j] = (idx >= first) ? desc_load(desc + idx) : desc_pack(DESC_PREFIX, 0);
            }
#pragma unroll
            for (int j = 0; j < PER_LANE; j++) {
                const unsigned inv = __ballot_sync(FULL, desc_state(d[j]) == DESC_INVALID);
                const unsigned pfx = __ballot_sync(FULL, desc_state(d[j]) == DESC_PREFIX);
                if (pfx) {
                    const unsigned nearer = (1u << (__ffs(pfx) - 1)) - 1u;
                    if (inv & nearer) pending = true;
                    break;
                }
                if (inv) {
                    pending = true;
                    break;
                }
            }
            if (pending) {
                __nanosleep(64);
                if (*dead) return 0;
                if (++spins > DECF_SPIN_LIMIT) {
                    if (lane == 0) {
                        *dead = 1;
                        atomicOr(err, (uint32_t)ERR_TIMEOUT);
                    }
                    return 0;
                }
            }
        } while (pending);
        bool done = false;
#pragma unroll
        for (int j = 0; j < PER_LANE; j++) {
            if (!done) {
                const unsigned pm = __ballot_sync(FULL, desc_state(d[j]) == DESC_PREFIX);
                const unsigned upto = pm ? (unsigned)(__ffs(pm) - 1) : 31u;
                excl += warp_sum64(lane <= upto ? desc_value(d[j]) : 0ull);
                if (pm) done = true;
            }
        }
        if (done) break;
        pos -= 32 * PER_LANE;
    }
    return excl;
}

__device__ __forceinline__ void decf_fetch_ticket(DecfInfo *slot, const uint4 *__restrict__ seq_rec, uint32_t *counter, uint32_t nseq)
{
    const uint32_t tk = atomicAdd(counter, 1u);
    DecfInfo I;
    I.seq = tk;
    I.b = 0;
    I.bytes = 0;
    I.pad = 0;
    I.src_off = 0;
    if (tk < nseq) {
        const uint4 r = __ldg(seq_rec + tk);
        I.src_off = ((uint64_t)r.y << 32) | r.x;
        I.b = r.z;
        I.bytes = r.w;
    }
    *slot = I;
}

__global__ void __launch_bounds__(DECF_CTA, 3)
k_dec_fused(const uint8_t *__restrict__ payload, DecPlan plan, DecTablesGlobal g, const uint4 *__restrict__ seq_rec,
            uint8_t *__restrict__ out, uint32_t *__restrict__ desc_end, uint64_t *__restrict__ desc_cnt,
            uint32_t *__restrict__ counter, uint32_t nseq, uint32_t *__restrict__ err, unsigned long long *__restrict__ trace)
{
    extern __shared__ __align__(16) unsigned char smem_raw[];
    DecfShared &sh = *reinterpret_cast<DecfShared *>(smem_raw);
    const unsigned tid = threadIdx.x, lane = lane_id(), warp = warp_id();
    volatile uint32_t *dead = &sh.dead;
    // SHAFA_FUSED_TRACE: CTA 0 stamps the phases of its first 128 sequences (decode thread 0: events 0-15, chain lane 0: 16-31)
    auto stamp = [&](uint32_t k, uint32_t ev) {
        if (trace && blockIdx.x == 0 && k < 128u) {
            unsigned long long tns;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(tns));
            trace[k * 32u + ev] = tns;
        }
    };

    if (tid == 0) {
        mbar_init(&sh.mb_payload, 1);
        mbar_init(&sh.mb_rawfree, 1);
        mbar_init(&sh.mb_carry, 1);
        mbar_init(&sh.mb_total, 1);
        mbar_init(&sh.mb_prefix, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        sh.dead = 0;
        // the first ticket and its payload
        decf_fetch_ticket(&sh.info[0], seq_rec, counter, nseq);
        if (sh.info[0].seq < nseq)
            bulk_load(sh.raw, payload + sh.info[0].src_off, sh.info[0].bytes, &sh.mb_payload);
        else
            mbar_arrive(&sh.mb_payload);
    }
    __syncthreads();

    if (warp == DEC_THREADS / 32) {
        // =========================== chain warp ===========================
        for (uint32_t k = 0;; k++) {
            const uint32_t par = k & 1u;
            const DecfInfo I = sh.info[par];
            if (I.seq >= nseq || *dead) break;
            if (lane == 0) stamp(k, 16);
            const uint64_t first_seq = plan.seq0[I.b];
            bool need_carry = (I.seq != first_seq), need_next = true;
            if (!need_carry && lane == 0) {
                sh.carry = 0;
                mbar_arrive(&sh.mb_carry);
            }
            // the predecessor's end offset and the next ticket, whichever can be had first
            uint32_t spins = 0;
            while (need_carry || need_next) {
                uint32_t got_carry = 0, got_next = 0;
                if (lane == 0) {
                    if (need_carry) {
                        const uint32_t v = ld_relaxed_u32(desc_end + I.seq - 1);
                        if (v >> 31) {
                            sh.carry = v & 0xffu;
                            mbar_arrive(&sh.mb_carry);
                            got_carry = 1;
                            stamp(k, 17);
                        }
                    }
                    if (need_next && mbar_test(&sh.mb_rawfree, par)) {
                        decf_fetch_ticket(&sh.info[par ^ 1u], seq_rec, counter, nseq);
                        const DecfInfo &N = sh.info[par ^ 1u];
                        if (N.seq < nseq)
                            bulk_load(sh.raw, payload + N.src_off, N.bytes, &sh.mb_payload);
                        else
                            mbar_arrive(&sh.mb_payload);
                        got_next = 1;
                        stamp(k, 18);
                    }
                }
                if (__shfl_sync(FULL, got_carry, 0)) need_carry = false;
                if (__shfl_sync(FULL, got_next, 0)) need_next = false;
                if (need_carry || need_next) {
                    __nanosleep(32);
                    if (*dead) break;
                    if (++spins > DECF_SPIN_LIMIT) {
                        if (lane == 0) {
                            *dead = 1;
                            atomicOr(err, (uint32_t)ERR_TIMEOUT);
                        }
                        break;
                    }
                }
            }
            if (*dead) break;
            // the sequence's symbol count -> its place in the block (decoupled look-back), for the deferred copy-out
            mbar_wait(&sh.mb_total, par, dead, err);
            if (lane == 0) stamp(k, 19);
            const uint32_t total = sh.total;
            uint64_t pfx = 0;
            if (I.seq != first_seq) {
                pfx = decf_lookback(desc_cnt, (int64_t)I.seq, (int64_t)first_seq, dead, err);
                if (lane == 0) desc_store(desc_cnt + I.seq, desc_pack(DESC_PREFIX, pfx + total));
            }
            if (lane == 0) {
                sh.prefix[par] = pfx;
                mbar_arrive(&sh.mb_prefix);
                stamp(k, 20);
            }
            __syncwarp();
        }
        return;
    }

    // =========================== decode warps ===========================
    DecfViewI t;
    t.pairs_s = smem_u32(sh.pairs);
    t.lutp_s = smem_u32(sh.lutp);
    t.slot0 = smem_u32(sh.slots) + tid * 4u;
    t.head0 = smem_u32(sh.heads) + tid;
    asm volatile("" : "+r"(t.pairs_s), "+r"(t.lutp_s), "+r"(t.slot0), "+r"(t.head0));  // not rebuilt from the shared window in the loops
    t.sleft = sh.sleft;
    t.slen = sh.slen;
    t.ssym = sh.ssym;
    const uint32_t stage_s = smem_opaque(sh.stage);
    const uint32_t row = tid * DEC_ROW;

    // the sequence whose symbols sit in the stage, waiting for their place in the output
    bool pend = false;
    uint32_t pend_total = 0, pend_par = 0, pend_k = 0;
    uint64_t pend_n = 0;
    uint8_t *pend_out = nullptr;
    bool pend_last = false;
    auto flush_pending = [&]() {
        // (called by all decode threads) look-back result of the pending sequence -> copy its symbols out
        mbar_wait(&sh.mb_prefix, pend_par, dead, err);
        if (tid == 0) stamp(pend_k + 1, 10);
        const uint64_t sfirst = sh.prefix[pend_par];
        if (pend_last && tid == 0 && sfirst + pend_total < pend_n) atomicOr(err, (uint32_t)ERR_PAYLOAD_SHORT);
        if (sfirst < pend_n) {  // else: only padding garbage
            const uint64_t room = pend_n - sfirst;
            const uint32_t syms = room < pend_total ? (uint32_t)room : pend_total;
            copy_out_unaligned<DEC_THREADS>(pend_out + sfirst, sh.stage, syms);
        }
        pend = false;
    };

    uint32_t cur_block = 0xffffffffu;
    for (uint32_t k = 0;; k++) {
        const uint32_t par = k & 1u;
        if (tid == 0) stamp(k, 0);
        mbar_wait(&sh.mb_payload, par, dead, err);  // also orders the chain warp's write of info[par]
        const DecfInfo I = sh.info[par];
        if (I.seq >= nseq || *dead) break;
        if (tid == 0) stamp(k, 1);
        const uint32_t seq = I.seq, b = I.b;
        const uint64_t first_seq = plan.seq0[b], z = plan.z[b], n = plan.n[b];
        const uint64_t sib = seq - first_seq;
        const bool is_last = (seq + 1 == plan.seq0[b + 1]);
        if (b != cur_block) {  // (the previous iteration's barriers: nobody uses the old tables any more)
            cur_block = b;
            const uint4 *src = reinterpret_cast<const uint4 *>(g.lutp + (uint64_t)b * DEC_LUT_SIZE);
            uint4 *dst = reinterpret_cast<uint4 *>(sh.lutp);
#pragma unroll
            for (int i = 0; i < DEC_LUT_SIZE * 4 / 16 / DEC_THREADS; i++) dst[i * DEC_THREADS + tid] = __ldg(src + i * DEC_THREADS + tid);
            sh.sleft[tid] = g.sleft[(uint64_t)b * 256 + tid];
            sh.slen[tid] = g.slen[(uint64_t)b * 256 + tid];
            sh.ssym[tid] = g.ssym[(uint64_t)b * 256 + tid];
        }
        // ---- payload of this sequence: landing zone -> rows ----
        {
            RowRegs rr;
            rr.v = *reinterpret_cast<const uint4 *>(sh.raw + tid * 16);
            rr.u = *reinterpret_cast<const uint2 *>(sh.raw + tid * 16 + 16);
            dec_store_row(sh.pairs, rr, z, sib);
        }
        dbar();  // rows and tables complete; the landing zone is free again
        if (tid == 0) {
            mbar_arrive(&sh.mb_rawfree);
            stamp(k, 2);
        }
        const uint64_t bits_left = (z - sib * DEC_SEQ_BYTES) * 8;
        const uint32_t limit = bits_left < (uint64_t)DEC_SEQ_BYTES * 8 ? (uint32_t)bits_left : DEC_SEQ_BYTES * 8;
        const uint32_t sub_lo = tid * DEC_SUB_BITS;
        const uint32_t stop = limit > sub_lo ? (limit - sub_lo < (uint32_t)DEC_SUB_BITS ? limit - sub_lo : (uint32_t)DEC_SUB_BITS) : 0u;

        // ---- pass 1: speculative decode from bit 0, symbols kept as records ----
        DecfSub u;
        u.mn = decf_pass(t, row, t.slot0, 0u, stop);
        u.start = 0;
        u.hn = 0;
        u.jsp = 0;
        u.count = u.mn.k;
        u.end = u.mn.end;
        sh.ends[tid] = (uint8_t)decf_over(u.end);
        if (tid == 0) stamp(k, 3);
        // has the predecessor's end arrived already?  Then the rounds run with the true start from the beginning.
        if (tid == 0) sh.carry_early = mbar_test(&sh.mb_carry, par) ? 1u : 0u;
        bool carry_known = false;
        uint32_t end0 = 0xffffffffu;  // end published before the true start was known (thread 255)
        for (int stage_no = 0; stage_no < 2; stage_no++) {
            uint32_t carry = 0;
            bool first_round = true;
            for (;;) {  // rounds until every sub-sequence starts where its predecessor ended
                dbar();
                if (first_round) {
                    first_round = false;
                    if (stage_no == 1 || sh.carry_early) {
                        carry = sh.carry;
                        carry_known = true;
                    }
                }
                const uint32_t want = tid ? (uint32_t)sh.ends[tid - 1] : carry;
                const bool changed = (want != u.start);
                if (!dbar_or(changed)) break;
                if (changed) {
                    decf_restart(t, row, want, stop, u);
                    sh.ends[tid] = (uint8_t)decf_over(u.end);
                }
            }
            if (tid == DEC_THREADS - 1) {
                const uint32_t e = decf_over(u.end);
                if (end0 == 0xffffffffu) {
                    st_relaxed_u32(desc_end + seq, 0x80000000u | e);
                    end0 = e;
                } else if (e != end0) {
                    atomicOr(err, (uint32_t)ERR_SPECULATION);  // the guess "the end does not depend on the start" was wrong
                }
            }
            if (tid == 0) stamp(k, 4 + 3 * stage_no);
            if (stage_no == 0) {
                // the previous sequence's symbols leave the stage here: between the speculative rounds and the true start
                if (pend) flush_pending();
                if (tid == 0) stamp(k, 5);
                if (carry_known) break;
                mbar_wait(&sh.mb_carry, par, dead, err);
                if (tid == 0) stamp(k, 6);
            }
        }

        // ---- symbol counts: CTA scan; the count goes to the chain warp, which finds the sequence's place ----
        uint32_t total;
        const uint32_t excl = decode_excl_scan(u.count | (u.mn.norec << 20), sh.scan, &total);  // (its barriers: the flush above is over)
        const bool direct = (total >> 20) != 0 || (total & 0xfffffu) > (uint32_t)DECF_STAGE;
        total &= 0xfffffu;
        const uint32_t my_excl = excl & 0xfffffu;
        if (tid == 0) stamp(k, 8);
        if (tid == 0) {
            if (seq != first_seq) lookback_publish(desc_cnt, (int64_t)seq, (int64_t)first_seq, total);
            else desc_store(desc_cnt + seq, desc_pack(DESC_PREFIX, total));
            sh.total = total;
            mbar_arrive(&sh.mb_total);
        }
        uint8_t *const out_block = out + plan.out_off[b];
        if (!direct) {
            // compaction at sequence-relative offsets
            uint32_t d = stage_s + my_excl;
            for (uint32_t i = 0; i < u.hn; i++) sts_u8(d + i, lds_u8(t.head0 + i * 256u));
            d += u.hn;
            const uint32_t mrec = u.mn.norec ? 0u : u.mn.m;
            for (uint32_t j = u.jsp; j < mrec; j++) {
                const uint32_t r = t.ld32s(t.slot0 + j * 1024u);
                sts_u8(d, r);
                if (r >> 29) sts_u8(d + 1, r >> 8);
                d += r >> 28;
            }
            pend = true;
            pend_total = total;
            pend_par = par;
            pend_k = k;
            pend_n = n;
            pend_out = out_block;
            pend_last = is_last;
            dbar();  // the stage is complete (the next sequence's flush reads it) and the rows are free
            if (tid == 0) stamp(k, 9);
        } else {
            // more symbols than the stage holds, or record lists that overflowed (codes of 1-2 bits): wait for the place
            // now and decode once more, straight to global memory
            mbar_wait(&sh.mb_prefix, par, dead, err);
            const uint64_t sfirst = sh.prefix[par];
            if (is_last && tid == 0 && sfirst + total < n) atomicOr(err, (uint32_t)ERR_PAYLOAD_SHORT);
            if (sfirst < n) {
                const uint64_t seq_room = n - sfirst;
                const uint32_t seq_syms = seq_room < total ? (uint32_t)seq_room : total;
                const uint32_t room = my_excl < seq_syms ? seq_syms - my_excl : 0u;
                uint8_t *dst = out_block + sfirst + my_excl;
                uint32_t q = u.start, i = 0, bad = 0;
                while (q < stop) {
                    const uint32_t e = decf_lookup(t, row, q, &bad);
                    if (i < room) dst[i] = (uint8_t)e;
                    i++;
                    q += decf_n(e) == 2u ? decf_len0(e) : decf_total(e);
                }
            }
            dbar();  // rows are free
        }
    }
    if (pend) flush_pending();
}

int launch_sf_decode_fused(shafa_cuda_ctx *ctx, const uint8_t *d_payload, const DecPlan &plan, const DecTablesGlobal &g,
                           const DecSeqRec *seq_rec, uint64_t nseq, uint8_t *d_sym, uint32_t *d_err)
{
    if (nseq == 0) return SHAFA_SUCCESS;
    const size_t smem = sizeof(DecfShared);
    int &occ = ctx->occ[OCC_DEC_FUSED];
    if (!occ) {
        SHAFA_CUDA_TRY(ctx, cudaFuncSetAttribute(k_dec_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        SHAFA_CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_dec_fused, DECF_CTA, smem));
        if (occ < 1) occ = 1;
    }
    // descriptors: [cnt u64 nseq][end u32 nseq]
    uint8_t *desc = (uint8_t *)scratch_get(ctx, SCR_DESC, nseq * 12 + 64);
    uint32_t *counter = misc_ptr(ctx, MISC_COUNTER);
    if (!desc || !counter) return SHAFA_LACK_OF_MEMORY;
    SHAFA_CUDA_TRY(ctx, cudaMemsetAsync(desc, 0, nseq * 12 + 64, ctx->stream));
    SHAFA_CUDA_TRY(ctx, cudaMemsetAsync(counter, 0, 4, ctx->stream));
    unsigned long long *trace = nullptr;
    const bool want_trace = getenv("SHAFA_FUSED_TRACE") != nullptr;
    if (want_trace) {
        trace = (unsigned long long *)scratch_get(ctx, SCR_DEC_FB, 128 * 32 * 8);
        if (trace) cudaMemsetAsync(trace, 0, 128 * 32 * 8, ctx->stream);
    }
    uint64_t grid = (uint64_t)ctx->sm_count * occ;
    {
        const char *e = getenv("SHAFA_FUSED_CTAS");  // experiments: resident CTAs per SM
        if (e && atoi(e) > 0 && atoi(e) < occ) grid = (uint64_t)ctx->sm_count * atoi(e);
    }
    if (grid > nseq) grid = nseq;
    SHAFA_LAUNCH(ctx, "dec_fused", k_dec_fused, dim3((unsigned)grid), dim3(DECF_CTA), smem, d_payload, plan, g, seq_rec, d_sym,
                 (uint32_t *)(desc + nseq * 8), (uint64_t *)desc, counter, (uint32_t)nseq, d_err, trace);
    SHAFA_CHECK_LAUNCH(ctx);
    if (want_trace && trace) {  // debugging aid: phase timestamps of CTA 0 to stderr (ns since its first stamp)
        static unsigned long long h[128 * 32];
        cudaStreamSynchronize(ctx->stream);
        cudaMemcpy(h, trace, sizeof h, cudaMemcpyDeviceToHost);
        const unsigned long long t0 = h[0];
        for (int k = 0; k < 128 && h[k * 32]; k++) {
            fprintf(stderr, "trace k=%d:", k);
            for (int e = 0; e < 32; e++)
                if (h[k * 32 + e]) fprintf(stderr, " e%d=%lld", e, (long long)(h[k * 32 + e] - t0));
            fprintf(stderr, "\n");
        }
    }
    return SHAFA_SUCCESS;
}

}  // namespace shafa
