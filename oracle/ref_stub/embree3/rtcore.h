// oracle/ref_stub/embree3/rtcore.h -- TEST INFRASTRUCTURE ONLY.
// Stand-in for Intel Embree 3's <embree3/rtcore.h> (absent from this image and from /root/reference: CMakeLists.txt:6 finds it on
// the build machine), just wide enough for the reference's OWN raytracing.cpp and light_probe.cpp to compile UNMODIFIED
// (oracle/Makefile `ref`).  Every ray query is answered by the CPU oracle's tracer (oracle/bvh.c: prt_o_closest_hit /
// prt_o_any_hit, the pinned ray/triangle rule of oracle/arith.h), so what these builds pin is everything the reference itself
// decides around the ray queries -- scene set-up, sampling, path logic, estimator, weights -- not Embree's own arithmetic.
// Conventions kept from Embree 3: rtcIntersect1 shortens ray.tfar to the hit distance and fills Ng (unnormalised (v1-v0)x(v2-v0)),
// primID, geomID; rtcOccluded1 sets ray.tfar = -inf on a hit; RTC_INVALID_GEOMETRY_ID marks a miss.
#pragma once
#include <cstddef>
#include <limits>

#define RTC_INVALID_GEOMETRY_ID ((unsigned int)-1)
#define RTC_MAX_INSTANCE_LEVEL_COUNT 1

enum RTCError { RTC_ERROR_NONE = 0, RTC_ERROR_UNKNOWN = 1, RTC_ERROR_INVALID_ARGUMENT = 2, RTC_ERROR_INVALID_OPERATION = 3, RTC_ERROR_OUT_OF_MEMORY = 4 };
enum RTCGeometryType { RTC_GEOMETRY_TYPE_TRIANGLE = 0 };
enum RTCBufferType { RTC_BUFFER_TYPE_INDEX = 0, RTC_BUFFER_TYPE_VERTEX = 1 };
enum RTCFormat { RTC_FORMAT_UINT3 = 0x5003, RTC_FORMAT_FLOAT3 = 0x9003 };

typedef struct RTCDeviceTy *RTCDevice;
typedef struct RTCSceneTy *RTCScene;
typedef struct RTCGeometryTy *RTCGeometry;
typedef void (*RTCErrorFunction)(void *userPtr, enum RTCError code, const char *str);

struct RTCRay {
    float org_x, org_y, org_z, tnear;
    float dir_x, dir_y, dir_z, time;
    float tfar;
    unsigned int mask, id, flags;
};
struct RTCHit {
    float Ng_x, Ng_y, Ng_z;
    float u, v;
    unsigned int primID, geomID;
    unsigned int instID[RTC_MAX_INSTANCE_LEVEL_COUNT];
};
struct RTCRayHit { struct RTCRay ray; struct RTCHit hit; };
struct RTCIntersectContext { unsigned int flags; void *filter; unsigned int instID[RTC_MAX_INSTANCE_LEVEL_COUNT]; };
inline void rtcInitIntersectContext(struct RTCIntersectContext *c) { c->flags = 0; c->filter = nullptr; c->instID[0] = RTC_INVALID_GEOMETRY_ID; }

RTCDevice rtcNewDevice(const char *config);
RTCError rtcGetDeviceError(RTCDevice);
void rtcSetDeviceErrorFunction(RTCDevice, RTCErrorFunction, void *userPtr);
RTCScene rtcNewScene(RTCDevice);
void rtcReleaseScene(RTCScene);
void rtcCommitScene(RTCScene);
RTCGeometry rtcNewGeometry(RTCDevice, enum RTCGeometryType);
void *rtcSetNewGeometryBuffer(RTCGeometry, enum RTCBufferType, unsigned int slot, enum RTCFormat, size_t byteStride, size_t itemCount);
void rtcCommitGeometry(RTCGeometry);
unsigned int rtcAttachGeometry(RTCScene, RTCGeometry);
void rtcReleaseGeometry(RTCGeometry);
void rtcIntersect1(RTCScene, struct RTCIntersectContext *, struct RTCRayHit *);
void rtcOccluded1(RTCScene, struct RTCIntersectContext *, struct RTCRay *);
