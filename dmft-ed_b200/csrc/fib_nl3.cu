// fib_nl3.cu -- fiber kernels for stars of 3 levels (Nbath = 2), full tiles; see hxv_fiber.cu / fiber_kernels.cuh
#include "fiber_kernels.cuh"
int fib_launch_nl3(int pass, cudaStream_t st, const FibArgs &A, int grid) { return fib_launch<3, false>(pass, st, A, grid); }
