// star_info.h -- star-product description of one spin basis (shared by hxv_star.cu and hxv_fiber.cu).
#pragma once
#include "edgpu_internal.h"
#include <vector>

static constexpr int kMaxBlocks = 4096;
static constexpr int kBigBlock = 2048;        // up-blocks at least this large use the pipelined kernel
static constexpr int kDotSlots = 16384;       // capacity of ctx->d_dotpart (per-CTA partial sums of the fused <x, H x>)
static constexpr int kBulkMin = 256;          // blocks at least this large use the copy-engine (TMA) kernels; smaller ones the fringe kernels

struct StarBlock {             // one occupation tuple of one spin
    int off;                   // first internal index
    int size;                  // prod D[n_a]
    int n[EDGPU_MAXORB];       // star occupations
    int sgn_lower[EDGPU_MAXORB];   // (-1)^{sum_{a'<a} n_a'} as 0/1
    uint32_t magic0;               // ceil(2^32 / D0): tid / D0 == __umulhi(tid, magic0) for tid < 1024
    int ny;                        // kNT / D0
    int nouter;                    // size / D0
    int nh;                        // largest hop count of any configuration of any star of this block
};

struct StarInfo {
    int norb = 0, nbath = 0, H = 0, ncfg = 0;
    int D[16], A0[16], coff[16];                 // per occupation m: star dim, #imp=0 configs, offset into cfg tables
    std::vector<StarBlock> blocks;
    // device copies
    StarBlock *d_blocks = nullptr;
    uint8_t *d_hopj = nullptr;                   // [ncfg][H]   target index inside the same-occupation star list
    int16_t *d_hopd = nullptr;                   // [ncfg][H]   target index minus own index (what the tiled kernels use)
    uint8_t *d_hopc = nullptr;                   // [ncfg]
    double *d_hopv = nullptr;                    // [norb][ncfg][H] signed amplitudes V_{a,k} * (-1)^{popc(bath below k)}
    double *d_estar = nullptr;                   // [norb][ncfg]    star diagonal energies
    double pair_e = 0.0;                         // (Ust-Jh): same-spin inter-orbital density term
    // tile schedule
    int *d_upgroups = nullptr;                   // [ngroups][2] = (start in d_uplist, count): groups of SMALL blocks
    int *d_uplist = nullptr;                     // block ids of the groups
    int ngroups = 0, max_small = 0;
    std::vector<int> big_blocks;                 // blocks handled by the persistent pipelined up kernel
    int max_block = 0;
    // "fringe": the configurations of all blocks smaller than kBulkMin, handled by one thread-per-element launch
    int nfringe = 0;
    int *d_fringe = nullptr;                     // [nfringe] internal indices
    ~StarInfo()
    {
        cudaFree(d_blocks); cudaFree(d_hopj); cudaFree(d_hopd); cudaFree(d_hopc); cudaFree(d_hopv); cudaFree(d_estar); cudaFree(d_upgroups); cudaFree(d_uplist); cudaFree(d_fringe);
    }
};

