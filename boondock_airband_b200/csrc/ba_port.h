/*
 * ba_port.h — the three spellings that differ between the product build (nvcc, sm_100a) and the development-time
 * thread emulation used by tests/emu (g++ -DBA_EMU, see tests/emu/cuda_emu.h; never part of libba_cuda.so).
 */
#ifndef BA_PORT_H
#define BA_PORT_H

#ifdef BA_EMU
#include "cuda_emu.h"
#define BA_SHARED(name) unsigned char* name = emu::tls().cta->smem
#define BA_LAUNCH(kern, grid, block, smem, stream, ...) \
    emu::launch(dim3(grid), dim3(block), (size_t)(smem), [=]() { kern(__VA_ARGS__); })
#define BA_BAR_SYNC(id, count) emu::bar_named((id), (count))
/* cp.async: global -> shared without passing through registers; synchronous in the emulation */
#define BA_CP_ASYNC_8(smem_ptr, gmem_ptr) memcpy((smem_ptr), (gmem_ptr), 8)
#define BA_CP_ASYNC_4(smem_ptr, gmem_ptr) memcpy((smem_ptr), (gmem_ptr), 4)
#define BA_CP_ASYNC_COMMIT() ((void)0)
#define BA_CP_ASYNC_WAIT(n) ((void)0)
#else
#include <cuda_runtime.h>
#define BA_SHARED(name) extern __shared__ __align__(16) unsigned char name[]
#define BA_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define BA_BAR_SYNC(id, count) asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory")
#define BA_CP_ASYNC_8(smem_ptr, gmem_ptr) \
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"((unsigned)__cvta_generic_to_shared(smem_ptr)), "l"(gmem_ptr) : "memory")
#define BA_CP_ASYNC_4(smem_ptr, gmem_ptr) \
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"((unsigned)__cvta_generic_to_shared(smem_ptr)), "l"(gmem_ptr) : "memory")
#define BA_CP_ASYNC_COMMIT() asm volatile("cp.async.commit_group;" ::: "memory")
#define BA_CP_ASYNC_WAIT(n) asm volatile("cp.async.wait_group %0;" ::"n"(n) : "memory")
#endif

#endif
