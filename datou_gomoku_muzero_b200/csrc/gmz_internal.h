// gmz_internal.h -- declarations shared by the translation units of libgmz.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>

#include "../../include/gmz.h"
#include "gmz_common.cuh"

struct gmz_engine {
    unsigned magic;          // GMZ_ENGINE_MAGIC while alive (gmz_destroy is idempotent per handle value)
    gmz_config cfg;
    Params p;
    int NC;
    int device;              // CUDA device the engine was created on
    size_t bytes;
    void *workspace;
};
#define GMZ_ENGINE_MAGIC 0x474d5a32u

// error text behind gmz_last_error() (thread-local, defined in gmz_engine.cu); both return 1
int gmz_fail(const char *fmt, const char *a = "");
int gmz_check_launch(const char *what);
extern "C" void gmz_set_error_(const char *msg);

// ticketed persistent play kernel, one translation unit per (MuZero mode, float32 accumulation) pair
struct PlayArgs;
int gmz_launch_play_mz0_f0(gmz_engine *e, const PlayArgs &a, cudaStream_t st);
int gmz_launch_play_mz0_f1(gmz_engine *e, const PlayArgs &a, cudaStream_t st);
int gmz_launch_play_mz1_f0(gmz_engine *e, const PlayArgs &a, cudaStream_t st);
int gmz_launch_play_mz1_f1(gmz_engine *e, const PlayArgs &a, cudaStream_t st);
