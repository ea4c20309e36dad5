/** \file tma.cuh
 * \brief 1D bulk asynchronous copies (TMA: cp.async.bulk -> SASS UBLKCP) completing on an mbarrier
 * (SASS SYNCS), CTA-local.  Used by csrstream.cu (tiles of scalar rows) and factor.cu (runs of
 * 5 x 5 blocks).  Source addresses and sizes must be multiples of 16 bytes: callers round the source
 * down and the size up (every device array is padded by 64 bytes at allocation, common.cuh).
 */
#ifndef B200_TMA_CUH
#define B200_TMA_CUH

namespace b200 {

__device__ __forceinline__ unsigned smem_u32(const void *p)
{
	return (unsigned)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(unsigned long long *bar, int count)
{
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count));
	asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes)
{
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
	             :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes,
                                         unsigned long long *bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity)
{
	asm volatile(
		"{\n"
		".reg .pred p;\n"
		"WAIT_LOOP:\n"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
		"@p bra WAIT_DONE;\n"
		"bra WAIT_LOOP;\n"
		"WAIT_DONE:\n"
		"}\n" :: "r"(smem_u32(bar)), "r"(parity) : "memory");
}

}  // namespace b200
#endif
