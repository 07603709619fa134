// export_pipe.h -- internal interface of the pipelined KmerSet export (export_pipe.cu, export_expand.cpp).
//
// dbg_export_kmerset hands the reference's consumer a P-slot table image (kmerSet.h:88-99) that is about half empty
// slots.  PCIe, not the GPU, bounds that hand-over, so the pipe ships only the OCCUPIED nodes (compacted on the device in
// slot order) plus the occupancy bitmap, and host threads expand them into the caller's array while later chunks are still
// on the link; when the host threads fall behind, a chunk travels as the plain image instead (DMA straight into the array).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace dbg {

struct ExportPipe;

// result codes: 0 ok, 1 = not applicable (table too small, no host threads, no memory): caller uses the plain copy,
// < 0 = CUDA error (message in err)
int export_pipe_run(ExportPipe **pipe, int device, cudaStream_t stream, const void *d_img, const uint32_t *d_nul32, uint64_t nul_words,
                    uint64_t P, int node_bytes, void *array, uint8_t *nul_flag, float *ms, uint64_t stats[4], char *err, int err_len);
void export_pipe_destroy(ExportPipe *pipe);

// host: array[s] = (bit s of bits, MSB first) ? next node of src : 0, for n_slots slots (n_slots % 8 == 0 or the last byte
// is partial); returns the number of nodes consumed.  node_bytes 16 or 32.  dst needs 16-byte alignment.
uint64_t expand_nodes(const uint8_t *bits, uint64_t n_slots, const void *src, void *dst, int node_bytes);

}  // namespace dbg
