// fib_nl7.cu -- fiber kernels for stars of 7 levels (Nbath = 6), full tiles; see hxv_fiber.cu / fiber_kernels.cuh
#include "fiber_kernels.cuh"
int fib_launch_nl7(int pass, cudaStream_t st, const FibArgs &A, int grid) { return fib_launch<7, false>(pass, st, A, grid); }
