// common.cuh — encodings and grid primitives shared by every kernel of the hot path.
// Reference citations are relative to the reference repository root.
#pragma once
#include "flatland_b200.h"

#include <cuda_runtime.h>

#define DEVI __device__ __forceinline__

namespace {

enum : int { WAITING = 0, READY = 1, MAL_OFF = 2, MOVING = 3, STOPPED = 4, MALFUNCTION = 5, DONE = 6 };  // states.py:5-12
enum : int { A_NOTHING = 0, A_LEFT = 1, A_FORWARD = 2, A_RIGHT = 3, A_STOP = 4 };                          // rail_env_action.py:5-10
constexpr int NPRED = FL_PRED_DEPTH + 1;  // prediction rows 0..500 (treeobs.cpp:50-65)

DEVI bool on_map(int s) { return s >= MOVING && s <= MALFUNCTION; }
DEVI bool off_map(int s) { return s <= MAL_OFF; }
DEVI int d_row(int d) { return (d == 2) - (d == 0); }
DEVI int d_col(int d) { return (d == 1) - (d == 3); }
// grid4.py:66-87 / tool.h:337-352: 4-bit nibble of orientation o, bit order N,E,S,W msb->lsb
DEVI int nibble(unsigned cell, int o) { return (cell >> ((3 - o) * 4)) & 0xF; }
DEVI int tbit(int nb, int d) { return (nb >> (3 - d)) & 1; }
DEVI int first_dir(int nb) { return __clz(nb) - 28; }  // first set bit in N,E,S,W order, nb in 1..15


}  // namespace
