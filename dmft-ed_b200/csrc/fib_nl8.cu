// fib_nl8.cu -- fiber kernels for stars of 8 levels (Nbath = 7), full tiles; see hxv_fiber.cu / fiber_kernels.cuh
#include "fiber_kernels.cuh"
int fib_launch_nl8(int pass, cudaStream_t st, const FibArgs &A, int grid) { return fib_launch<8, false>(pass, st, A, grid); }
