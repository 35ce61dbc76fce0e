// Recording stand-in for <GL/gl.h>: just the OpenGL 1.1 client-array entry points nerf::NeRF::DrawCPUMesh uses.
// TEST ONLY (tests/host/draw_check.cpp defines the functions and logs the calls).
#pragma once
typedef unsigned int GLenum;
typedef int GLint;
typedef int GLsizei;
typedef void GLvoid;
#define GL_TRIANGLES 0x0004
#define GL_UNSIGNED_BYTE 0x1401
#define GL_UNSIGNED_INT 0x1405
#define GL_FLOAT 0x1406
#define GL_VERTEX_ARRAY 0x8074
#define GL_NORMAL_ARRAY 0x8075
#define GL_COLOR_ARRAY 0x8076
extern "C" {
void glEnableClientState(GLenum array);
void glDisableClientState(GLenum array);
void glVertexPointer(GLint size, GLenum type, GLsizei stride, const GLvoid* pointer);
void glNormalPointer(GLenum type, GLsizei stride, const GLvoid* pointer);
void glColorPointer(GLint size, GLenum type, GLsizei stride, const GLvoid* pointer);
void glDrawElements(GLenum mode, GLsizei count, GLenum type, const GLvoid* indices);
}
