/*
 * ilqr_nccl_dyn.h -- NCCL bound at run time (dlopen), so that neither libilqr_cuda.so nor the model plug-ins carry
 * a link-time dependency on it: single-GPU users never load it, and a host process that already holds an NCCL
 * (PyTorch bundles one under the same SONAME) shares that copy.  Only the five entry points the final gather of
 * SURVEY.md section 8e needs; prototypes restated from nccl.h (NCCL 2.x ABI).
 */
#pragma once
#include <dlfcn.h>
#include <stddef.h>

#include <initializer_list>

namespace ilqr_nccl {

typedef struct ncclComm* comm_t;
typedef struct { char internal[128]; } unique_id; /* NCCL_UNIQUE_ID_BYTES */
enum { DT_CHAR = 0 };                             /* ncclInt8 / ncclChar */

struct Api {
    int (*GetUniqueId)(unique_id*) = nullptr;
    int (*CommInitRank)(comm_t*, int, unique_id, int) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, comm_t, void* /*cudaStream_t*/) = nullptr;
    int (*CommDestroy)(comm_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    void* lib = nullptr;
};

/* returns nullptr (and the dlerror text in *why) when NCCL cannot be loaded */
inline const Api* api(const char** why) {
    static Api a;
    static bool tried = false;
    static const char* err = nullptr;
    if (!tried) {
        tried = true;
        for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
            a.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (a.lib) break;
        }
        if (!a.lib) {
            err = "libnccl.so.2 not found (dlopen)";
        } else {
            a.GetUniqueId = (int (*)(unique_id*))dlsym(a.lib, "ncclGetUniqueId");
            a.CommInitRank = (int (*)(comm_t*, int, unique_id, int))dlsym(a.lib, "ncclCommInitRank");
            a.AllGather = (int (*)(const void*, void*, size_t, int, comm_t, void*))dlsym(a.lib, "ncclAllGather");
            a.CommDestroy = (int (*)(comm_t))dlsym(a.lib, "ncclCommDestroy");
            a.GetErrorString = (const char* (*)(int))dlsym(a.lib, "ncclGetErrorString");
            if (!a.GetUniqueId || !a.CommInitRank || !a.AllGather || !a.CommDestroy || !a.GetErrorString) err = "NCCL symbols missing";
        }
    }
    if (err) { if (why) *why = err; return nullptr; }
    return &a;
}

} /* namespace ilqr_nccl */
