// Shadows CUDA's cuda_gl_interop.h (which needs <GL/gl.h>): the one entry point nerf_model.cu's mesh upload names.
#pragma once
#include <cuda_runtime.h>
#include <GL/glew.h>
extern "C" cudaError_t cudaGraphicsGLRegisterBuffer(struct cudaGraphicsResource** resource, GLuint buffer, unsigned int flags);
