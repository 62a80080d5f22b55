// Product kernels: structured fast collision, FMA contraction allowed.
#define LBX_COLLIDE CollideFast
#define LBX_GETTER launchers_fast
#include "kernels_impl.inc"
