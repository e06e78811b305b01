// oracle/ref_embree_stub.cpp -- TEST INFRASTRUCTURE ONLY.  The Embree 3 calls of the reference (ref_stub/embree3/rtcore.h) answered by
// the CPU oracle's tracer, and the App singleton of ref_stub/platform/platform.h.
#include <embree3/rtcore.h>
#include "platform/platform.h"
#include "prt_oracle.h"
#include <cmath>
#include <cstdlib>
#include <vector>

struct RTCDeviceTy { RTCErrorFunction fn = nullptr; void *user = nullptr; };
struct RTCGeometryTy { std::vector<unsigned char> verts, idx; size_t vstride = 12, istride = 12, nv = 0, nt = 0; int refs = 1; };
struct RTCSceneTy { RTCGeometryTy *geom = nullptr; prt_o_scene *sc = nullptr; };

RTCDevice rtcNewDevice(const char *) { return new RTCDeviceTy(); }
RTCError rtcGetDeviceError(RTCDevice) { return RTC_ERROR_NONE; }
void rtcSetDeviceErrorFunction(RTCDevice d, RTCErrorFunction fn, void *user) { if (d) { d->fn = fn; d->user = user; } }
RTCScene rtcNewScene(RTCDevice) { return new RTCSceneTy(); }
void rtcReleaseScene(RTCScene s) {
    if (!s) return;
    if (s->sc) prt_o_scene_destroy(s->sc);
    if (s->geom && --s->geom->refs == 0) delete s->geom;
    delete s;
}
RTCGeometry rtcNewGeometry(RTCDevice, enum RTCGeometryType) { return new RTCGeometryTy(); }
void *rtcSetNewGeometryBuffer(RTCGeometry g, enum RTCBufferType type, unsigned int, enum RTCFormat, size_t stride, size_t count) {
    std::vector<unsigned char> &b = type == RTC_BUFFER_TYPE_VERTEX ? g->verts : g->idx;
    b.assign(stride * count + 16, 0);
    if (type == RTC_BUFFER_TYPE_VERTEX) { g->vstride = stride; g->nv = count; } else { g->istride = stride; g->nt = count; }
    return b.data();
}
void rtcCommitGeometry(RTCGeometry) {}
unsigned int rtcAttachGeometry(RTCScene s, RTCGeometry g) { s->geom = g; g->refs++; return 0; }
void rtcReleaseGeometry(RTCGeometry g) { if (g && --g->refs == 0) delete g; }
void rtcCommitScene(RTCScene s) {
    if (!s->geom) return;
    std::vector<uint32_t> tri(3 * s->geom->nt);
    for (size_t i = 0; i < s->geom->nt; i++)
        for (int k = 0; k < 3; k++) tri[3 * i + k] = *(const uint32_t *)(s->geom->idx.data() + i * s->geom->istride + 4 * k);
    s->sc = prt_o_scene_create((const float *)s->geom->verts.data(), s->geom->vstride, (uint32_t)s->geom->nv, tri.data(), (uint32_t)s->geom->nt);
}
void rtcIntersect1(RTCScene s, struct RTCIntersectContext *, struct RTCRayHit *rh) {
    float o[3] = { rh->ray.org_x, rh->ray.org_y, rh->ray.org_z }, d[3] = { rh->ray.dir_x, rh->ray.dir_y, rh->ray.dir_z }, t, ng[3];
    uint32_t prim;
    if (s->sc && prt_o_closest_hit(s->sc, o, d, rh->ray.tnear, rh->ray.tfar, 1, &t, &prim, ng)) {
        rh->ray.tfar = t;
        rh->hit.Ng_x = ng[0]; rh->hit.Ng_y = ng[1]; rh->hit.Ng_z = ng[2];
        rh->hit.primID = prim; rh->hit.geomID = 0; rh->hit.u = rh->hit.v = 0.f;
    }
}
void rtcOccluded1(RTCScene s, struct RTCIntersectContext *, struct RTCRay *r) {
    float o[3] = { r->org_x, r->org_y, r->org_z }, d[3] = { r->dir_x, r->dir_y, r->dir_z };
    if (s->sc && prt_o_any_hit(s->sc, o, d, r->tnear, r->tfar, 1)) r->tfar = -INFINITY;
}

App &App::get() { static App app; return app; }
