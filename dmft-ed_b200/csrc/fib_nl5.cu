// fib_nl5.cu -- fiber kernels for stars of 5 levels (Nbath = 4), full tiles; see hxv_fiber.cu / fiber_kernels.cuh
#include "fiber_kernels.cuh"
int fib_launch_nl5(int pass, cudaStream_t st, const FibArgs &A, int grid) { return fib_launch<5, false>(pass, st, A, grid); }
