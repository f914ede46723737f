// fib_nl8h.cu -- fiber kernels for stars of 8 levels (Nbath = 7), half tiles; see hxv_fiber.cu / fiber_kernels.cuh
#include "fiber_kernels.cuh"
int fib_launch_nl8h(int pass, cudaStream_t st, const FibArgs &A, int grid) { return fib_launch<8, true>(pass, st, A, grid); }
