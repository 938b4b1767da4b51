// Entry points declared in afsync.h whose kernels are not written yet.  Each one
// fails loudly (no CPU fallback).  This file shrinks as kernels land.
#include "afs_common.cuh"

#define AFS_PENDING(name) afs::fail(AFS_ERR_UNSUPPORTED, name ": kernel not implemented yet")

extern "C" {
int afs_wtw_create(afs_wtw **, int, const double *, const int64_t *, const int64_t *, int, int, int) { return AFS_PENDING("afs_wtw_create"); }
int afs_wtw_destroy(afs_wtw *) { return AFS_OK; }
int afs_wtw_state_bytes(const afs_wtw *, size_t *) { return AFS_PENDING("afs_wtw_state_bytes"); }
int afs_wtw_reset(afs_wtw *, void *, void *) { return AFS_PENDING("afs_wtw_reset"); }
int afs_wtw_push(afs_wtw *, const double *, int, const uint8_t *, int32_t *, void *) { return AFS_PENDING("afs_wtw_push"); }
int afs_wtw_path_layout(const afs_wtw *, int, int64_t *, int64_t *) { return AFS_PENDING("afs_wtw_path_layout"); }
int afs_wtw_path_ptr(const afs_wtw *, const int32_t **, const int32_t **) { return AFS_PENDING("afs_wtw_path_ptr"); }
int afs_wtw_positions_ptr(const afs_wtw *, const int32_t **) { return AFS_PENDING("afs_wtw_positions_ptr"); }
}
