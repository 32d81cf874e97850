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
#else
#include <cuda_runtime.h>
#define BA_SHARED(name) extern __shared__ __align__(16) unsigned char name[]
#define BA_LAUNCH(kern, grid, block, smem, stream, ...) kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#define BA_BAR_SYNC(id, count) asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory")
#endif

#endif
