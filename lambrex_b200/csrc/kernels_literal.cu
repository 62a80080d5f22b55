// Parity kernels: the reference's operation order; this file is compiled with
// --fmad=false so results are bit-identical to a non-FMA CPU build.
#define LBX_COLLIDE CollideLiteral
#define LBX_GETTER launchers_literal
#include "kernels_impl.inc"
