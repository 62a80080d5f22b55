// -*- mode: c++ -*-
// Process-wide start-up / shut-down, same two calls as /root/reference/include/lambrex.h:6-7.
// lambrexInit selects the GPU ($LOCAL_RANK or 0), creates the library's stream and fails
// loudly (amrex::Abort) when no CUDA device is present: there is no CPU path.
#ifndef LBX_LAMBREX_H
#define LBX_LAMBREX_H
void lambrexInit();
void lambrexFinalise();
#endif
