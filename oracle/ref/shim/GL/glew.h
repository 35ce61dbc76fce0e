// Stand-in for the OpenGL names of RO-MAP's mesh upload path (MeshData VBO handles, nerf_model.cu:2100-2180).
// TEST INFRASTRUCTURE ONLY: declarations so that nerf_model.cu compiles; the harness never calls the mesh path and
// defines these symbols as aborting stubs.
#pragma once
typedef unsigned int GLuint;
typedef unsigned int GLenum;
typedef int GLsizei;
typedef long GLsizeiptr;
#define GL_ARRAY_BUFFER 0x8892
#define GL_ELEMENT_ARRAY_BUFFER 0x8893
#define GL_STATIC_DRAW 0x88E4
extern "C" {
void glGenBuffers(GLsizei n, GLuint* buffers);
void glBindBuffer(GLenum target, GLuint buffer);
void glBufferData(GLenum target, GLsizeiptr size, const void* data, GLenum usage);
void glDeleteBuffers(GLsizei n, const GLuint* buffers);
}
