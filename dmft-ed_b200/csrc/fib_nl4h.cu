// fib_nl4h.cu -- fiber kernels for stars of 4 levels (Nbath = 3), half tiles; see hxv_fiber.cu / fiber_kernels.cuh
#include "fiber_kernels.cuh"
int fib_launch_nl4h(int pass, cudaStream_t st, const FibArgs &A, int grid) { return fib_launch<4, true>(pass, st, A, grid); }
