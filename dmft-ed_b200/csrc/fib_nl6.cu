// fib_nl6.cu -- fiber kernels for stars of 6 levels (Nbath = 5), full tiles; see hxv_fiber.cu / fiber_kernels.cuh
#include "fiber_kernels.cuh"
int fib_launch_nl6(int pass, cudaStream_t st, const FibArgs &A, int grid) { return fib_launch<6, false>(pass, st, A, grid); }
