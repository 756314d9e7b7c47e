/*
 * chunk_extension.h -- backend-defined types for the B200 CUDA backend.
 *
 * Drop this directory into the reference tree as TeaLeaf/c_kernels/cuda/ and build with
 * `make KERNELS=cuda` (see Makefile here and INTEGRATION.md).  TeaLeaf/chunk.h:7 includes this
 * header; chunk.h:39-66 declares the chunk's fields as FieldBufferType and chunk.h:78 / chunk.c:29
 * allocate one ChunkExtension per chunk.  Reference counterpart:
 * TeaLeaf/c_kernels/sycl/chunk_extension.h:6-13.
 *
 * The host never dereferences a FieldBufferType: here it is a small handle
 * naming one of the backend's HBM-resident fields.
 */
#pragma once

struct tl_chunk;
struct tl_comms;

typedef struct TlFieldRef
{
    int id; /* TL_FIELD_* of include/tealeaf_b200.h */
} TlFieldRef;

typedef TlFieldRef* FieldBufferType;

typedef struct ChunkExtension
{
    struct tl_chunk* handle; /* the GPU-resident chunk (libtealeaf_b200.so) */
    TlFieldRef refs[12];
} ChunkExtension;
