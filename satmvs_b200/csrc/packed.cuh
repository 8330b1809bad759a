// packed.cuh — Blackwell packed fp32 arithmetic (fma/mul/add/sub.rn.f32x2 -> FFMA2/FMUL2/FADD2): two
// IEEE-rounded fp32 results per instruction, operands are 64-bit register pairs.
#pragma once
namespace satmvs {
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float lo, float hi) { u64 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ void upk(u64 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) { u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
}  // namespace satmvs
