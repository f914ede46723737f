// fib_nl4.cu -- fiber kernels for stars of 4 levels (Nbath = 3), full tiles; see hxv_fiber.cu / fiber_kernels.cuh
#include "fiber_kernels.cuh"
int fib_launch_nl4(int pass, cudaStream_t st, const FibArgs &A, int grid) { return fib_launch<4, false>(pass, st, A, grid); }
