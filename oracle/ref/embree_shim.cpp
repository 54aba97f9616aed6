// embree_shim.cpp — the Embree 3 entry points the reference calls, on the oracle's BVH. TEST INFRASTRUCTURE ONLY.
// See shim/embree3/rtcore.h. (Shared buffers are read at rtcCommitScene, like Embree does for a static scene.)
#include "../oracle.h"

#include <embree3/rtcore.h>

#include <atomic>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <vector>

// rtcIntersect1 calls, per thread (summed by ref_shim_intersect_calls): the metric counts every traversal query,
// the reference's own counter (_total_rays) only path segments
namespace
{
    std::mutex                                 g_mu;
    std::vector<std::atomic<unsigned long long> *> g_counters;
    struct Counter
    {
        std::atomic<unsigned long long> n { 0 };
        Counter()
        {
            std::lock_guard<std::mutex> lk(g_mu);
            g_counters.push_back(&n);
        }
    };
    thread_local Counter *t_counter = nullptr;
}    // namespace
extern "C" unsigned long long ref_shim_intersect_calls()
{
    std::lock_guard<std::mutex> lk(g_mu);
    unsigned long long          s = 0;
    for (auto *c : g_counters) s += c->load(std::memory_order_relaxed);
    return s;
}

struct RTCDeviceTy
{
    int unused;
};
struct RTCGeometryTy
{
    const char *verts = nullptr, *index = nullptr, *attr = nullptr;
    size_t      vstride = 0, istride = 0, astride = 0, nverts = 0, ntris = 0, nattr = 0;
};
struct RTCSceneTy
{
    RTCGeometryTy     *geom = nullptr;
    orc_rawbvh        *bvh  = nullptr;
    std::vector<float> soup;    // 9 floats per triangle (v0, v1, v2), gathered through the index buffer
};

extern "C" {
RTCDevice   rtcNewDevice(const char *) { return new RTCDeviceTy(); }
RTCScene    rtcNewScene(RTCDevice) { return new RTCSceneTy(); }
RTCGeometry rtcNewGeometry(RTCDevice, enum RTCGeometryType) { return new RTCGeometryTy(); }
void rtcSetSharedGeometryBuffer(RTCGeometry g, enum RTCBufferType type, unsigned int, enum RTCFormat, const void *ptr, size_t off, size_t stride, size_t count)
{
    const char *p = static_cast<const char *>(ptr) + off;
    if (type == RTC_BUFFER_TYPE_VERTEX) g->verts = p, g->vstride = stride, g->nverts = count;
    if (type == RTC_BUFFER_TYPE_INDEX) g->index = p, g->istride = stride, g->ntris = count;
    if (type == RTC_BUFFER_TYPE_VERTEX_ATTRIBUTE) g->attr = p, g->astride = stride, g->nattr = count;
}
void     rtcSetGeometryVertexAttributeCount(RTCGeometry, unsigned int) {}
void     rtcCommitGeometry(RTCGeometry) {}
unsigned rtcAttachGeometry(RTCScene s, RTCGeometry g)
{
    s->geom = g;
    return 0;
}
void rtcCommitScene(RTCScene s)
{
    RTCGeometryTy *g = s->geom;
    s->soup.resize(g->ntris * 9);
    for (size_t t = 0; t < g->ntris; t++)
    {
        const uint32_t *idx = reinterpret_cast<const uint32_t *>(g->index + t * g->istride);
        for (int k = 0; k < 3; k++) std::memcpy(&s->soup[t * 9 + size_t(k) * 3], g->verts + size_t(idx[k]) * g->vstride, 12);
    }
    if (s->bvh) orc_rawbvh_destroy(s->bvh);
    s->bvh = orc_rawbvh_create(s->soup.data(), uint32_t(g->ntris));
}
void rtcIntersect1(RTCScene s, struct RTCIntersectContext *, struct RTCRayHit *rh)
{
    if (!t_counter) t_counter = new Counter();    // (leaked with the thread: a handful per process)
    t_counter->n.store(t_counter->n.load(std::memory_order_relaxed) + 1, std::memory_order_relaxed);
    const float o[3] = { rh->ray.org_x, rh->ray.org_y, rh->ray.org_z }, d[3] = { rh->ray.dir_x, rh->ray.dir_y, rh->ray.dir_z };
    float       t, u, v;
    uint32_t    prim;
    if (!s->bvh || !orc_rawbvh_intersect(s->bvh, o, d, rh->ray.tnear, rh->ray.tfar, &t, &u, &v, &prim)) return;
    const float *p  = &s->soup[size_t(prim) * 9];
    const float  e1[3] = { p[3] - p[0], p[4] - p[1], p[5] - p[2] }, e2[3] = { p[6] - p[0], p[7] - p[1], p[8] - p[2] };
    rh->ray.tfar = t;
    rh->hit.u = u, rh->hit.v = v;
    rh->hit.Ng_x   = e1[1] * e2[2] - e2[1] * e1[2];    // (v1-v0) x (v2-v0), same operation order as glm::cross
    rh->hit.Ng_y   = e1[2] * e2[0] - e2[2] * e1[0];
    rh->hit.Ng_z   = e1[0] * e2[1] - e2[0] * e1[1];
    rh->hit.primID = prim, rh->hit.geomID = 0, rh->hit.instID[0] = RTC_INVALID_GEOMETRY_ID;
}
void rtcInterpolate0(RTCGeometry g, unsigned int primID, float u, float v, enum RTCBufferType, unsigned int, float *P, unsigned int n)
{
    const uint32_t *idx = reinterpret_cast<const uint32_t *>(g->index + size_t(primID) * g->istride);
    const float    *a = reinterpret_cast<const float *>(g->attr + size_t(idx[0]) * g->astride), *b = reinterpret_cast<const float *>(g->attr + size_t(idx[1]) * g->astride),
                   *c = reinterpret_cast<const float *>(g->attr + size_t(idx[2]) * g->astride);
    const float w = 1.0f - u - v;
    for (unsigned int i = 0; i < n; i++) P[i] = (w * a[i] + u * b[i]) + v * c[i];
}
}
