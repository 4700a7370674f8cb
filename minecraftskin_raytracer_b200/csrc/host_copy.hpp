// host_copy.hpp — a few host threads that move finished pieces of a frame from the library's page-locked staging image to
// a caller's pageable image (an Image's std::vector<Color>: what TileRenderer::render returns, tile_renderer.cpp:129-189)
// while the GPU is still shading the rest of the frame.
#pragma once
#include <cuda_runtime.h>

#include <cstddef>
#include <vector>

namespace mcskin {

struct HostCopyJob {
    cudaEvent_t after;  // the piece is in the staging image once this event has completed (null: it already is)
    unsigned char* dst;
    const unsigned char* src;
    size_t dstPitch, srcPitch, rowBytes, rows;
};

// Runs the jobs on the pool's threads and on the calling thread; returns when all are done.
// false: waiting for an event failed (the CUDA error is left for the caller to report).
bool run_host_copies(int device, const std::vector<HostCopyJob>& jobs);

// (exposed for the CPU tests) rows of rowBytes from src to dst; streaming stores when both are 16-byte aligned
void copy_rows_streaming(unsigned char* dst, const unsigned char* src, size_t dstPitch, size_t srcPitch, size_t rowBytes,
                         size_t rows);

}  // namespace mcskin
