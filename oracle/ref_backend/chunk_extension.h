/*
 * chunk_extension.h for the ORACLE backend (test infrastructure only).
 * Supplies the two backend-defined types that TeaLeaf/chunk.h:7,39-66,78 needs
 * (reference counterpart: TeaLeaf/c_kernels/sycl/chunk_extension.h:6-13).
 * Fields are plain host arrays.
 */
#pragma once
typedef double* FieldBufferType;

typedef struct ChunkExtension
{
    int unused;
} ChunkExtension;
