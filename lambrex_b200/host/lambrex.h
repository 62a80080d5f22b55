// -*- mode: c++ -*-
// Process-wide start-up / shut-down, same two calls as /root/reference/include/lambrex.h:6-7.
// lambrexInit selects the GPU ($LOCAL_RANK or 0), creates the library's stream and fails
// loudly (amrex::Abort) when no CUDA device is present: there is no CPU path.
#ifndef LBX_LAMBREX_H
#define LBX_LAMBREX_H
#include <cstddef>

#include "AmrSim.h"   // as /root/reference/include/lambrex.h:4: one include gives callers the whole API
void lambrexInit();
void lambrexFinalise();
// Addition: distributed start-up, one process per GPU of one NVSwitch box (the role MPI_Init plays
// under amrex::Initialize).  `allgather(send, bytes, recv, user)` gathers `bytes` from every rank
// in rank order and returns 0 -- host plumbing only (torch.distributed, MPI ...).  Boxes of every
// level are then owned by ranks (amrex::DistributionMapping) and neighbours' boxes are read through
// CUDA-IPC peer pointers.  Collective.
void lambrexInitParallel(int rank, int nranks, int (*allgather)(const void*, size_t, void*, void*), void* user);
// testing aid: the box-ownership view of THIS process (nranks = 1: every box local, no collectives)
void lambrexSetParallelView(int rank, int nranks);
#endif
