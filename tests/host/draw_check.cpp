// draw_check.cpp — nerf::NeRF::DrawCPUMesh compiled against a recording OpenGL stand-in (tests/host/gl_stub): prints the
// call sequence the viewer's GL context would receive.  Built together with ro_map_b200/host/nerf_host.cpp; no GPU needed.
#include <cstdio>
#include <thread>

#include <GL/gl.h>

#include "nerf.h"

static nerf::CPUMeshData* g_mesh = nullptr;
static const char* which(const void* p) {
    if (p == g_mesh->verts.data()) return "verts";
    if (p == g_mesh->normals.data()) return "normals";
    if (p == g_mesh->colors.data()) return "colors";
    if (p == g_mesh->indices.data()) return "indices";
    return "?";
}
extern "C" {
void glEnableClientState(GLenum a) { printf("enable %x\n", a); }
void glDisableClientState(GLenum a) { printf("disable %x\n", a); }
void glVertexPointer(GLint n, GLenum t, GLsizei s, const GLvoid* p) { printf("vertex %d %x %d %s\n", n, t, s, which(p)); }
void glNormalPointer(GLenum t, GLsizei s, const GLvoid* p) { printf("normal %x %d %s\n", t, s, which(p)); }
void glColorPointer(GLint n, GLenum t, GLsizei s, const GLvoid* p) { printf("color %d %x %d %s\n", n, t, s, which(p)); }
void glDrawElements(GLenum m, GLsizei c, GLenum t, const GLvoid* p) { printf("draw %x %d %x %s\n", m, c, t, which(p)); }
}

int main() {
    nerf::NeRF::GPUnum = 1;   // the constructor only assigns ids and a GPU slot
    nerf::NeRF obj;
    nerf::CPUMeshData& m = obj.GetCPUMeshData();
    g_mesh = &m;
    printf("-- no mesh yet\n");
    obj.DrawCPUMesh();
    m.verts = {0, 0, 0, 1, 0, 0, 0, 1, 0};
    m.normals = {0, 0, 1, 0, 0, 1, 0, 0, 1};
    m.colors = {255, 0, 0, 0, 255, 0, 0, 0, 255};
    m.indices = {0, 1, 2};
    m.have_reslult = true;
    printf("-- mesh present\n");
    obj.DrawCPUMesh();
    printf("-- mesh being updated by the training thread\n");
    {
        std::unique_lock<std::mutex> busy(m.mesh_mutex);
        std::thread viewer([&] { obj.DrawCPUMesh(); });   // must return at once without drawing
        viewer.join();
    }
    printf("-- manager entry point\n");
    obj.DrawMesh();
    return 0;
}
