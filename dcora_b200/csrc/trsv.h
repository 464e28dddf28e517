// Block triangular-solve schedule shared by trsv_host.cpp (construction, host
// self check) and sptrsv.cu (upload + kernel).
#pragma once
#include "common.cuh"

namespace dcora {

constexpr int TRSV_SMALL_BLOCK = 32;  // diagonal blocks of <= 32 rows: one warp, lane = row
constexpr int TRSV_MAX_BLOCK = 512;   // larger ND blocks are split into a chain

// One level of the block elimination tree (one direction, one agent).
struct TrsvLevel {
  int ta0, ta1;  // phase-A row tasks
  int gpt;       // 4-lane groups cooperating on one phase-A task (1, 2, 4 or 8)
  int sb0, sb1;  // small diagonal blocks (phase B, warp per block)
  int lr0, lr1;  // rows of large diagonal blocks (phase B, warp per row)
  int pad;
};

struct AgentTrsvDir {
  // phase A: t_B = rhs_B - M[B, solved] * y
  std::vector<int> ta_row;    // factor row of the task (pose start); ~row: scalar row
  std::vector<int> ta_bptr;   // ntasks + 1: range of (d+1)x(d+1) blocks
  std::vector<int> bcol;      // factor row of the source pose block
  std::vector<double> bval;   // [a][b] = M(row + a, bcol + b)
  std::vector<int> srowptr;   // k + 1: scalar entries per factor row
  std::vector<int> scol;
  std::vector<double> sval;
  // phase B: y_B = D_B t_B  (forward) / x_B = D_B^T t_B (backward)
  std::vector<int> sb_row0, sb_w, sb_direct;
  std::vector<int64_t> sb_off;
  std::vector<int> lb_row0, lb_w, lb_direct;
  std::vector<int64_t> lb_off;
  std::vector<int> lr_blk, lr_i;
  std::vector<double> dense;  // packed triangular inverse blocks
  std::vector<TrsvLevel> levels;
};

struct AgentTrsv {
  int k = 0, nblocks = 0;
  AgentTrsvDir dir[2];  // 0: forward (L), 1: backward (L^T)
};

int build_agent_trsv(const HostFactor &F, int dh, AgentTrsv *out);
void agent_trsv_apply_host(const AgentTrsv &S, int dh, const double *b, double *x);

}  // namespace dcora
