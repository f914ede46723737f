// fib_nl6h.cu -- fiber kernels for stars of 6 levels (Nbath = 5), half tiles; see hxv_fiber.cu / fiber_kernels.cuh
#include "fiber_kernels.cuh"
int fib_launch_nl6h(int pass, cudaStream_t st, const FibArgs &A, int grid) { return fib_launch<6, true>(pass, st, A, grid); }
