// Block triangular-solve "tape" shared by trsv_host.cpp (construction, host
// self check) and sptrsv.cu (upload + kernel).
//
// The solve of one agent is compiled, at set-up, into one TAPE PER WARP of the
// agent's thread-block cluster: the exact sequence of work items that warp
// executes, with the factor values it needs laid out in consumption order.
// The kernel streams each tape through a per-warp shared-memory ring with bulk
// asynchronous copies (TMA), so every byte of the factor is read from HBM
// exactly once, sequentially, and -- being static data -- arbitrarily far ahead
// of the barriers; only the gathers from the solution panels depend on them.
//
// Everything is ONE kind of work item, a TILE TASK:
//     dst[r0 .. r0+nrows) (-)= sum_t  M_t . src[c_t .. c_t+DH)
// with DH x DH tiles M_t (DH = d+1; a pose is DH consecutive factor rows).  With the
// factor regrouped by nested-dissection blocks B (diagonal block L_BB, D_B = inv(L_BB))
// the solve  x = (L L^T)^-1 b  is rewritten so that ONE barrier per level of the block
// elimination tree suffices in each direction:
//     Lp[C, B] = L[C, B] D_B            (off-diagonal part with D folded in)
//     G_B      = D_B^T D_B              (= inv(L_BB L_BB^T), dense symmetric)
//   forward   T_C -= Lp[C, B] T_B       for the blocks B below C   (T starts as P b)
//   mid       X_B  = G_B T_B            any time after T_B is final
//   backward  X_B -= Lp[C, B]^T X_C     for the blocks C above B
// (T_B = L_BB y_B of the textbook forward solve; D_B T_B never has to be formed on the
// critical path.)  Three task types select the panels:
//     TAPE_FWD: src T, dst T, accumulating     TAPE_MID: src T, dst X, assigning
//     TAPE_BWD: src X, dst X, accumulating
// Accumulating tasks may be cut into pieces that run in different barrier slots: a tile may
// run in any slot after its SOURCE block is final (trsv_host.cpp uses that to take the long
// rows of the top separators off the critical path and to keep the warps evenly loaded).
//
// A tape is a sequence of fixed-size RECORDS of 16 + 8 DH^2 8-byte words (doubles, or two
// packed int32), one record per iteration of a round; the kernel moves one record per bulk
// copy (TMA) into one slot of the warp's shared-memory ring:
//   word 0..3    source rows of the record TAPE_AHEAD records further on (two int32 per
//                word, one per lane group): the kernel gathers its panel rows that many
//                iterations ahead and never has to wait for a future record to do so
//   word 4       {niter, gpt | type << 8 | last_of_slot << 16}   (first record of a round)
//   word 5..8    8 x int32 (dst_row0 << 3 | nrows), -1: none     (first record of a round)
//   word 9       {records from this one to the end of its barrier slot, 0}   (first record of a round)
//   word 10..11  padding
//   word 12..15  the record's own source rows (host replay; unused by the kernel)
//   word 16..    8 x DH x DH doubles, lane (group, b) owns column b of its tile.  DH = 4:
//                [half][lane][2] -- rows 0,1 of every lane's column, then rows 2,3 -- so
//                that each of the lane's two 128-bit shared loads is a conflict-free,
//                contiguous 512 bytes per warp; DH = 3: [group][b][a]
// Every tape starts with a round of TAPE_AHEAD zero records (they carry the source rows of
// the first real records); it belongs to the first barrier slot.
// A round = `niter` records executed by one warp: `gpt` adjacent groups share one task
// (their partial sums are added at the end of the round).  A cluster barrier follows the
// round flagged last_of_slot; every warp of the agent has exactly one such round per slot
// (a zero record when it has no work there).  Rows are GLOBAL panel rows (agent offset
// included).
#pragma once
#include "common.cuh"

namespace dcora {

constexpr int TRSV_MAX_BLOCK = 512;   // larger ND blocks are split into a chain
constexpr int TAPE_REC_HEADER = 16;   // words before the tile values of a record
constexpr int TAPE_AHEAD = 4;         // records between a record and the one whose rows it names
                                      // (= depth of the kernel's ring: its gather look-ahead)
inline int tape_record_words(int dh) { return TAPE_REC_HEADER + 8 * dh * dh; }
// word (behind the header) of tile element (row a, column b) of lane group g
inline int tape_value_word(int dh, int g, int b, int a) {
  if (dh == 4) return (a >> 1) * 64 + (g * 4 + b) * 2 + (a & 1);
  return (g * dh + b) * dh + a;
}

enum TapeType : int { TAPE_FWD = 0, TAPE_MID = 1, TAPE_BWD = 2 };

struct AgentTape {
  int k = 0, nblocks = 0, nwarps = 0, dh = 0;
  int nslots = 0;
  int nlev[2] = {0, 0};
  std::vector<std::vector<uint64_t>> warp_tape;  // nwarps tapes, whole records
  int64_t stored_values = 0;  // factor values on the tapes, both directions (incl. padding)
  // cost model of the schedule: sum over slots of the busiest warp's
  // iterations / rounds, and the perfectly balanced iteration count
  int64_t crit_iters = 0, crit_rounds = 0, ideal_iters = 0;
};

// `row_offset`: global panel row of the agent's first factor row (== its
// first internal column); the tape stores global rows / columns.
// The agent's work is dealt to `nwarps` warps; the tapes of warps
// [nwarps, nwarps_total) only take part in the barriers.
int build_agent_tape(const HostFactor &F, int dh, int nwarps, int nwarps_total, int row_offset,
                     AgentTape *out);

// Sequential host replay of the tapes (same grouping and operand pairing as
// the kernel): x = (L L^T)^-1 b with b, x in INTERNAL column order (agent
// local).  R = 1.  T, X: scratch panels of >= k doubles.
void agent_tape_apply_host(const AgentTape &S, const HostFactor &F, int row_offset,
                           const double *b, double *x, double *T, double *X);

}  // namespace dcora
