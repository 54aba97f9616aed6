/* embree3/rtcore.h — SHIM. TEST INFRASTRUCTURE ONLY (oracle/ref).
 *
 * The reference links Intel Embree 3.x (CMakeLists.txt:22 find_package(embree 3.0), readme.md:26 -> 3.13.0); it is not
 * vendored and not installable in this image. This header declares exactly the Embree 3 API subset the reference calls
 * (src/render/entities/components.h:67-69, src/objects/model.cpp:10-47,60-94); embree_shim.cpp implements it on top of the
 * oracle's own BVH + Moeller-Trumbore test (orc_rawbvh_*), with Embree's published conventions: Ng = (v1-v0)x(v2-v0)
 * unnormalised, barycentrics (u,v) weight v1,v2, hit iff tnear < t <= tfar, rtcInterpolate0 = (1-u-v)*a0 + u*a1 + v*a2.
 * The third-party arithmetic is RESTATED here, the reference's own sources above it are compiled unmodified. */
#pragma once
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct RTCDeviceTy   *RTCDevice;
typedef struct RTCSceneTy    *RTCScene;
typedef struct RTCGeometryTy *RTCGeometry;

#define RTC_INVALID_GEOMETRY_ID ((unsigned int) -1)
enum RTCGeometryType { RTC_GEOMETRY_TYPE_TRIANGLE = 0 };
enum RTCBufferType { RTC_BUFFER_TYPE_INDEX = 0, RTC_BUFFER_TYPE_VERTEX = 1, RTC_BUFFER_TYPE_VERTEX_ATTRIBUTE = 2 };
enum RTCFormat { RTC_FORMAT_UINT3 = 0x5003, RTC_FORMAT_FLOAT2 = 0x9002, RTC_FORMAT_FLOAT3 = 0x9003 };

struct RTCRay
{
    float        org_x, org_y, org_z, tnear;
    float        dir_x, dir_y, dir_z, time;
    float        tfar;
    unsigned int mask, id, flags;
};
struct RTCHit
{
    float        Ng_x, Ng_y, Ng_z;
    float        u, v;
    unsigned int primID, geomID, instID[1];
};
struct RTCRayHit
{
    struct RTCRay ray;
    struct RTCHit hit;
};
struct RTCIntersectContext
{
    int          flags;
    void        *filter;
    unsigned int instID[1];
};
static inline void rtcInitIntersectContext(struct RTCIntersectContext *c)
{
    c->flags = 0, c->filter = 0, c->instID[0] = RTC_INVALID_GEOMETRY_ID;
}

RTCDevice   rtcNewDevice(const char *config);
RTCScene    rtcNewScene(RTCDevice);
RTCGeometry rtcNewGeometry(RTCDevice, enum RTCGeometryType);
void        rtcSetSharedGeometryBuffer(RTCGeometry, enum RTCBufferType, unsigned int slot, enum RTCFormat, const void *ptr, size_t byteOffset, size_t byteStride,
                                       size_t itemCount);
void        rtcSetGeometryVertexAttributeCount(RTCGeometry, unsigned int);
void        rtcCommitGeometry(RTCGeometry);
unsigned    rtcAttachGeometry(RTCScene, RTCGeometry);
void        rtcCommitScene(RTCScene);
void        rtcIntersect1(RTCScene, struct RTCIntersectContext *, struct RTCRayHit *);
void        rtcInterpolate0(RTCGeometry, unsigned int primID, float u, float v, enum RTCBufferType, unsigned int slot, float *P, unsigned int valueCount);

#ifdef __cplusplus
}
#endif
