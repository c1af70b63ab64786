// sf_decode.cuh -- declarations shared by the two Shannon-Fano decoders: the fused single-pass kernel
// (sf_decode_fused.cu) and the exact multi-kernel path (sf_decode.cu).
#pragma once

#include "common.cuh"
#include "core.h"
#include "internal.h"

namespace shafa {

// per-block decode tables in global memory (built by k_dec_tables)
struct DecTablesGlobal {
    uint16_t *lut;    // [nb][4096]  first-level: (sym << 8) | len
    uint8_t *lut2;    // [nb][8192]  (nsym << 4) | bits of all whole codes in a 13-bit window
    uint32_t *lut3;   // [nb][4096]  first one or two codes of the window with their symbols
    uint32_t *lutp;   // [nb][4096]  pair LUT of the fused decoder (core.h: decf_entry)
    uint64_t *sleft;  // [nb][256]
    uint8_t *slen;    // [nb][256]
    uint8_t *ssym;    // [nb][256]
    uint8_t *flags;   // [nb]        DecBlockFlags
    uint8_t *lut16;   // [lean slots][65536]  all whole codes of a 16-bit window (core.h: dec_lut16_entry), lean blocks only
};

enum DecBlockFlags : uint8_t {
    DECB_COMPLETE = 1,  // the table is a complete prefix-free code: every bit pattern decodes
    DECB_NOSYNC = 2,    // nearly one code length: such codes do not self-synchronise
    DECB_WIDE = 4,      // a code longer than 32 bits (the fused decoder's end offsets are bytes)
    DECB_SHORT = 8      // no code longer than the lut1 window (12 bits)
};
// "lean" block: complete code of nearly one length without long codes, and enough payload to pay for a 64 KiB table.
// Its sequences are synchronised by k_dec_sync_lean (every start walked, several codes per lookup), everybody else's by
// k_dec_sync.
constexpr uint64_t DEC_LEAN_MIN_Z = 4ull * 65536;
__host__ __device__ inline bool dec_block_lean(uint8_t f, uint64_t z)
{
    return (f & (DECB_COMPLETE | DECB_NOSYNC | DECB_SHORT)) == (DECB_COMPLETE | DECB_NOSYNC | DECB_SHORT) && z >= DEC_LEAN_MIN_Z;
}
// a block goes through the fused kernel iff its table is complete, synchronising and not wide
__host__ __device__ inline bool dec_block_fused_ok(uint8_t f) { return (f & (DECB_COMPLETE | DECB_NOSYNC | DECB_WIDE)) == DECB_COMPLETE; }

struct DecPlan {
    const uint64_t *pay_off;  // [nb]
    const uint64_t *z;        // [nb]
    const uint64_t *n;        // [nb] symbols per block
    const uint64_t *out_off;  // [nb] output offset of each block
    const uint64_t *seq0;     // [nb+1] first sequence of each block
    uint32_t nb;
    // the sequences split into two classes, each numbered densely in block order (k_dec_classes):
    const uint64_t *cls_seq0[2];  // [nb+1] first class-c sequence number of each block (class 0: k_dec_sync, 1: k_dec_sync_lean)
    const uint32_t *lean_slot;    // [nb] which lut16 table a lean block uses
};

// what a ticket of the fused kernel needs before anything else: {payload offset lo, hi, block, bytes to copy}
typedef uint4 DecSeqRec;

// Thread s builds the row of sub-sequence s: payload bytes [16 s, 16 s + 24) of the sequence as six
// big-endian words; bytes at or past z read as zero.
struct RowRegs {
    uint4 v;
    uint2 u;
};
// the six big-endian words of this thread's sub-sequence (bytes at or past z read as zero)
__device__ __forceinline__ void dec_row_words(const RowRegs &r, uint64_t z, uint64_t seq_in_block, uint32_t (&w)[DEC_ROW_WORDS])
{
    const uint64_t o = seq_in_block * DEC_SEQ_BYTES + (uint64_t)threadIdx.x * 16;
    w[0] = r.v.x, w[1] = r.v.y, w[2] = r.v.z, w[3] = r.v.w, w[4] = r.u.x, w[5] = r.u.y;
    if (o + 4 * DEC_ROW_WORDS > z) {  // only at the end of a block
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
__device__ __forceinline__ void dec_store_row(uint64_t *pairs, const RowRegs &r, uint64_t z, uint64_t seq_in_block)
{
    const unsigned tid = threadIdx.x;
    uint32_t w[DEC_ROW_WORDS];
    dec_row_words(r, z, seq_in_block, w);
    uint64_t row[DEC_ROW];
    dec_row_from_words(w, row);
#pragma unroll
    for (int k = 0; k < DEC_ROW; k++) pairs[tid * DEC_ROW + k] = row[k];
}

int launch_sf_decode_fused(shafa_cuda_ctx *ctx, const uint8_t *d_payload, const DecPlan &plan, const DecTablesGlobal &g,
                           const DecSeqRec *seq_rec, uint64_t nseq, uint8_t *d_sym, uint32_t *d_err);

}  // namespace shafa
