// TEST INFRASTRUCTURE ONLY -- link-time stand-ins for the nine Embree 2.x entry points the
// reference's Scene.cpp references (Scene.cpp:198-211,360,416,466). Embree 2.7 ships with the
// reference only as macOS/Windows binaries, so the CPU intersection path cannot run here; the
// two intersect entry points abort to make any accidental use loud.
#include <embree2/rtcore.h>
#include <embree2/rtcore_ray.h>
#include <stdio.h>
#include <stdlib.h>

namespace {
	struct StubMesh { void* vertices; void* indices; };
	StubMesh* asMesh(RTCScene s) { return reinterpret_cast<StubMesh*>(s); }
}

RTCScene rtcNewScene(RTCSceneFlags, RTCAlgorithmFlags) {
	StubMesh* m = static_cast<StubMesh*>(calloc(1, sizeof(StubMesh)));
	return reinterpret_cast<RTCScene>(m);
}
unsigned rtcNewTriangleMesh(RTCScene s, RTCGeometryFlags, size_t triangles, size_t vertices, size_t) {
	asMesh(s)->vertices = malloc(vertices * 16 + 16);
	asMesh(s)->indices = malloc(triangles * 12 + 16);
	return 0;
}
void* rtcMapBuffer(RTCScene s, unsigned, RTCBufferType type) {
	return type == RTC_VERTEX_BUFFER ? asMesh(s)->vertices : asMesh(s)->indices;
}
void rtcUnmapBuffer(RTCScene, unsigned, RTCBufferType) {}
void rtcSetMask(RTCScene, unsigned, int) {}
void rtcCommit(RTCScene) {}
void rtcDeleteScene(RTCScene s) {
	free(asMesh(s)->vertices);
	free(asMesh(s)->indices);
	free(asMesh(s));
}
void rtcIntersect(RTCScene, RTCRay&) {
	fprintf(stderr, "racc ref shim: Embree is not available on this platform\n");
	abort();
}
void rtcIntersect8(const void*, RTCScene, RTCRay8&) {
	fprintf(stderr, "racc ref shim: Embree is not available on this platform\n");
	abort();
}
