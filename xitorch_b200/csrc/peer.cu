// Peer-visible device memory for the exchange regions of the row-sharded eigensolver (include/xitorch_b200.h,
// xt_peer_*): whole cudaMalloc allocations shared between the one-process-per-GPU ranks through CUDA IPC handles.
// New work: the reference has no multi-GPU code.
#include "common.cuh"

#include <cstring>

extern "C" {

int xt_peer_alloc(size_t bytes, void** ptr_out, void* handle_out) {
  XT_REQUIRE(bytes > 0 && ptr_out != nullptr && handle_out != nullptr, "peer_alloc: bad arguments");
  static_assert(sizeof(cudaIpcMemHandle_t) == XT_PEER_HANDLE_BYTES, "handle size");
  void* p = nullptr;
  XT_CUDA_OK(cudaMalloc(&p, bytes));
  cudaError_t e = cudaMemset(p, 0, bytes);
  if (e == cudaSuccess) e = cudaDeviceSynchronize();
  cudaIpcMemHandle_t h;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&h, p);
  if (e != cudaSuccess) {
    cudaFree(p);
    XT_CUDA_OK(e);
  }
  memcpy(handle_out, &h, sizeof(h));
  *ptr_out = p;
  return XT_OK;
}

int xt_peer_open(const void* handle, void** ptr_out) {
  XT_REQUIRE(handle != nullptr && ptr_out != nullptr, "peer_open: bad arguments");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle, sizeof(h));
  void* p = nullptr;
  XT_CUDA_OK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
  *ptr_out = p;
  return XT_OK;
}

int xt_peer_close(void* ptr) {
  if (ptr != nullptr) XT_CUDA_OK(cudaIpcCloseMemHandle(ptr));
  return XT_OK;
}

int xt_peer_free(void* ptr) {
  if (ptr != nullptr) XT_CUDA_OK(cudaFree(ptr));
  return XT_OK;
}

}  // extern "C"
