// Shared declarations for the jetb200 CUDA library (sm_100a only).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "jetb200.h"

namespace jb {

// ---- error plumbing: every C entry point returns int and leaves a thread-local message --------
std::string &LastError();
int Fail(const std::string &msg);

#define JB_CUDA(expr)                                                                             \
    do {                                                                                          \
        cudaError_t jb_err__ = (expr);                                                            \
        if (jb_err__ != cudaSuccess) {                                                            \
            return ::jb::Fail(std::string(#expr) + ": " + cudaGetErrorString(jb_err__));          \
        }                                                                                         \
    } while (0)

#define JB_TRY(expr)                                                                              \
    do {                                                                                          \
        int jb_rc__ = (expr);                                                                     \
        if (jb_rc__ != 0)                                                                         \
            return jb_rc__;                                                                       \
    } while (0)

#define JB_REQUIRE(cond, msg)                                                                     \
    do {                                                                                          \
        if (!(cond))                                                                              \
            return ::jb::Fail(msg);                                                               \
    } while (0)

inline size_t ElemBytes(int dtype) { return dtype == JB_C64 ? 8 : 16; }
inline bool IsPow2(int64_t x) { return x > 0 && (x & (x - 1)) == 0; }
inline int Log2(int64_t x)
{
    int l = 0;
    while ((int64_t(1) << l) < x)
        l++;
    return l;
}

int NumSMs();

// Resident CTAs per SM of a kernel (cached per function pointer): persistent grids are sized as
// NumSMs() * this so that every CTA of the grid is resident at once (no partial tail wave).
int BlocksPerSM(const void *kernel, int threads, size_t dyn_smem);
// cudaFuncAttributeMaxDynamicSharedMemorySize, set once per (device, kernel)
int EnsureDynamicSmem(const void *kernel, size_t bytes);
template <typename K> inline int PersistentBlocksPerSM(K kernel, int threads, size_t dyn_smem)
{
    return BlocksPerSM(reinterpret_cast<const void *>(kernel), threads, dyn_smem);
}

// ---- K1: permutation ---------------------------------------------------------------------------
// Bit-permutation description: the tensor has n address bits (all extents powers of two);
// output address bit j is input address bit src[j].
struct BitPerm {
    int n = 0;
    uint8_t src[64];
};

int LaunchPermute(int dtype, const void *in, void *out, int rank, const int64_t *extent,
                  const int32_t *perm, cudaStream_t stream);

// ---- K2: dense row-major GEMM -------------------------------------------------------------------
size_t GemmWorkspaceBytes(int dtype, int64_t m, int64_t n, int64_t k);
int LaunchGemm(int dtype, int64_t m, int64_t n, int64_t k, const void *a, const void *b, void *c,
               void *ws, size_t ws_bytes, cudaStream_t stream);

// tensor-core (tcgen05, 3xTF32) complex64 GEMM for large aligned shapes (gemm_tc.cu)
bool GemmDmmaEligible(int dtype, int64_t m, int64_t n, int64_t k); // complex128, FP64 tensor pipe (gemm_dmma.cu)
size_t GemmDmmaWorkspaceBytes(int64_t m, int64_t n, int64_t k);
int LaunchGemmDmmaGatherA(int64_t m, int64_t n, int64_t k, const void *a_tensor, const int *free_bits, int n_free,
                          const int *common_bits, int n_common, const void *b, void *c, void *ws, size_t ws_bytes,
                          cudaStream_t stream);
int LaunchGemmDmma(int64_t m, int64_t n, int64_t k, const void *a, const void *b, void *c, void *ws, size_t ws_bytes,
                   cudaStream_t stream);
bool GemmTcEligible(int dtype, int64_t m, int64_t n, int64_t k);
size_t GemmTcWorkspaceBytes(int64_t m, int64_t n, int64_t k);
int LaunchGemmTc(int64_t m, int64_t n, int64_t k, const void *a, const void *b, void *c, void *ws,
                 size_t ws_bytes, cudaStream_t stream);

// ---- fused contraction --------------------------------------------------------------------------
// A prepared pairwise contraction: everything derived from shapes/modes on the host once, so that
// launching it (many times, from a CUDA graph capture) is a pure kernel launch.
struct StreamParams; // defined in contract.cu

struct ContractPlan {
    int dtype = JB_C64;
    int rank_a = 0, rank_b = 0, rank_c = 0;
    std::vector<int64_t> extent_a, extent_b, extent_c;
    std::vector<int32_t> modes_a, modes_b, modes_c;
    int64_t m = 1, n = 1, k = 1;
    int kernel = 0; // 0 = stream, 1 = ttgt
    // ttgt: which GEMM kernel LaunchGemm will pick for this shape (JB_GEMM_*: what actually launches)
    int gemm_kind = 0;
    size_t ws_bytes = 0;
    // ttgt
    bool permute_a = false, permute_b = false;
    std::vector<int32_t> perm_a, perm_b;
    size_t ws_a_off = 0, ws_b_off = 0, ws_gemm_off = 0, ws_gemm_bytes = 0;
    // ttgt, complex64: C^T = B^T A^T — the large operand B plays the GEMM's A role (read as it is, full
    // 128-row tiles) and the small A is the one expanded to B'; the small C^T is transposed at the end
    bool swap_roles = false;
    size_t ws_ct_off = 0;
    // ttgt, complex128: the DMMA GEMM reads A in its original layout (no permuted copy of A)
    bool gather_a = false;
    std::vector<int> a_free_bits, a_common_bits; // address bits of A's free / contracted index bits, ascending
    // ttgt, DOTU / GEMV corner with power-of-two extents: both operands are read ONCE in their original layouts
    // (no permuted copies): blob = DotGatherParams (contract.cu)
    bool gather_dot = false;
    std::vector<unsigned char> dot_blob;
    // ttgt, small output (M, N <= 16) and a moderate K, power-of-two extents: one CTA per contraction reads both operands
    // in place and is batched over slices (SmallGemmGatherKernel); blob = SmallGemmParams (contract.cu)
    bool small_gemm = false;
    std::vector<unsigned char> small_blob;
    // stream kernel parameters (opaque blob, see contract.cu)
    std::vector<unsigned char> stream_blob;
    int launches = 1;
    double flops() const { return 8.0 * double(m) * double(n) * double(k); }
    double bytes() const
    {
        return double(ElemBytes(dtype)) *
               (double(m) * double(k) + double(k) * double(n) + double(m) * double(n));
    }
};

// The kernel LaunchGemm dispatches (dtype, m, n, k) to when the plan's workspace is provided: JB_GEMM_* codes.
int GemmKind(int dtype, int64_t m, int64_t n, int64_t k);
// Slice batching: one launch processes `count` slices whose per-slice tensors lie `stride` bytes apart (operands
// shared by all slices have stride 0).  count == 1 (or a null pointer) is the plain launch.
struct BatchArgs {
    int count = 1;
    long long stride_a = 0, stride_b = 0, stride_c = 0; // bytes between consecutive slices' A / B / C
};
struct ChainBatchArgs {
    int count = 1;
    long long stride_x0 = 0, stride_xk = 0;
    long long stride_r[16] = {0}; // per step: bytes between consecutive slices' small operand (0 = shared)
};

int MakeContractPlan(int dtype, int rank_a, const int64_t *extent_a, const int32_t *modes_a,
                     int rank_b, const int64_t *extent_b, const int32_t *modes_b,
                     ContractPlan *plan);
int LaunchContract(const ContractPlan &plan, const void *a, const void *b, void *c, void *ws,
                   cudaStream_t stream, const BatchArgs *batch = nullptr);

// ---- fused contraction chain (chain.cu) ---------------------------------------------------------
// A run of ContractTensors calls in which each result is contracted next with a small tensor,
// executed as ONE kernel that keeps the intermediates in shared memory.
struct ChainOperand {
    std::vector<int32_t> modes;
    std::vector<int64_t> extent;
    bool x_is_left = true; // the chained tensor is operand A of this ContractTensors call
};

struct ChainOp {
    int dtype = JB_C64;
    int n_steps = 0;
    int log_tile = 0;
    int conflict_free = 1;
    int n_stages = 0;
    int launches = 1; // chain kernel, plus the matrix gather when a register stage reads the constant bank
    int register_steps = 0; // steps executed inside register stages
    std::vector<unsigned char> blob; // ChainParams (chain_plan.h)
    std::vector<int32_t> modes_c;
    std::vector<int64_t> extent_c;
    double flops = 0.0;      // 8*M*N*K summed over the steps
    double step_bytes = 0.0; // sizeof(T)*(MK+KN+MN) summed over the steps (unfused traffic)
    double bytes = 0.0;      // sizeof(T)*(|X_0| + sum |R_i| + |X_k|): what the fused launch must move
};

int ChainMaxTileBits(int dtype);
bool ChainTilePaddingEnabled(); // false when JB_CHAIN_NO_PAD=1: short chains keep their minimal tiles
bool ChainFusionEnabled(); // false when JB_DISABLE_CHAIN=1 is set in the environment
bool ChainStepEligible(const ContractPlan &cp, bool *x_is_left);
// returns non-zero (reason in *why) when the chain does not fit one tile; not an error
int MakeChainOp(int dtype, const std::vector<int32_t> &modes_x, const std::vector<int64_t> &extent_x,
                const std::vector<ChainOperand> &ops, int max_tile_bits, ChainOp *out,
                std::string *why, bool allow_register_stages = true);
// The step matrices of a chain's register stages live in a constant-bank slot (a gather kernel writes
// them there before the chain kernel): plans acquire a slot for their lifetime (-1: none free -> no
// fusion), the operator-level entry points share ChainOperatorSlot().
int ChainAcquireSlot(int device);
void ChainReleaseSlot(int device, int slot);
int ChainOperatorSlot();
int LaunchChain(const ChainOp &op, const void *x0, const void *const *r, void *xk, int slot,
                cudaStream_t stream, const ChainBatchArgs *batch = nullptr);

// ---- elementwise ---------------------------------------------------------------------------------
int LaunchAdd(int dtype, int64_t n, const void *a, const void *b, void *c, cudaStream_t stream);
int LaunchConj(int dtype, int64_t n, const void *in, void *out, cudaStream_t stream);
int LaunchSlice(int dtype, const void *in, void *out, int rank, const int64_t *extent, int axis,
                int64_t value, cudaStream_t stream);

} // namespace jb
