// fib_nl7h.cu -- fiber kernels for stars of 7 levels (Nbath = 6), half tiles; see hxv_fiber.cu / fiber_kernels.cuh
#include "fiber_kernels.cuh"
int fib_launch_nl7h(int pass, cudaStream_t st, const FibArgs &A, int grid) { return fib_launch<7, true>(pass, st, A, grid); }
