// Block triangular-solve "tape" shared by trsv_host.cpp (construction, host
// self check) and sptrsv.cu (upload + kernel).
//
// The schedule of one agent is compiled, at set-up, into one TAPE PER WARP of
// the agent's thread-block cluster: the exact sequence of work items that warp
// executes over all levels of the forward and the backward walk, with the
// factor values it needs laid out in consumption order.  The kernel streams
// each tape through a per-warp shared-memory ring with bulk asynchronous
// copies (TMA), so every byte of the factor is read from HBM exactly once,
// sequentially, and -- being static data -- arbitrarily far ahead of the level
// barriers; only the gathers from the solution panel depend on a barrier.
//
// A tape is a sequence of 8-byte words (doubles, or two packed int32):
//   for dir in {forward, backward}: for level in 0..nlev[dir]-1:
//     A header   {n_rounds, a_present}        a_present: the level has a phase A
//     n_rounds x round:
//        {niter, gpt}
//        8 x slot ( {row0 | -1, nrows}, {perm_base, 0} )   one slot per 4-lane group
//        zero padding to a multiple of 4 words
//        niter x ( 4 words: source rows c[0..7] ; 8 x DH x DH doubles: [group][b][a],
//                  i.e. lane (group, b) owns DH consecutive words = column b of its block )
//     B header   {n_small, n_large}
//     n_small x  {row0, w} {direct, 0} ceil(w/2) words perm[] , w(w+1)/2 doubles
//     n_large x  {src_row0, len} {dest_row, dest_perm} , len doubles
// Phase A:  T[row0+a] = rhs[row0+a] - sum_blocks M[a][:] . Y[c .. c+DH)
// Phase B:  Y[B] = D_B T[B]  (forward: packed lower, backward: D_B^T)
// Rows are GLOBAL panel rows (agent offset included); perm entries are global
// internal columns.
#pragma once
#include "common.cuh"

namespace dcora {

constexpr int TRSV_SMALL_BLOCK = 32;  // diagonal blocks of <= 32 rows: one warp, lane = row
constexpr int TRSV_MAX_BLOCK = 512;   // larger ND blocks are split into a chain
constexpr int TAPE_CHUNK_WORDS = 256; // 2 KB bulk-copy granularity of the tape ring

struct AgentTape {
  int k = 0, nblocks = 0, nwarps = 0, dh = 0;
  int nlev[2] = {0, 0};
  std::vector<std::vector<uint64_t>> warp_tape;  // nwarps tapes, padded to whole chunks
  int64_t stored_values = 0;  // factor values on the tapes, both directions (incl. padding)
};

// `row_offset`: global panel row of the agent's first factor row (== its
// first internal column); the tape stores global rows / columns.
// The agent's work is dealt to `nwarps` warps; the tapes of warps
// [nwarps, nwarps_total) only take part in the barriers.
int build_agent_tape(const HostFactor &F, int dh, int nwarps, int nwarps_total, int row_offset,
                     AgentTape *out);

// Sequential host replay of the tapes (same grouping and operand pairing as
// the kernel): x = (L L^T)^-1 b with b, x in INTERNAL column order, global
// indexing (only the agent's rows are touched).  R = 1.
void agent_tape_apply_host(const AgentTape &S, const double *b, double *x,
                           double *Tpanel, double *Ypanel);

}  // namespace dcora
