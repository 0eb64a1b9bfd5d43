/* TEST INFRASTRUCTURE ONLY -- not product code.
 *
 * Minimal stand-in for <OpenCL/cl.h> so that the UNMODIFIED reference sources
 * (RayAccelerator/Scene.cpp, Bvh2.cpp, Environment.cpp, ...) compile on Linux.
 * The only behaviour implemented is "a buffer is a malloc'd copy of the host
 * pointer", which is enough to capture the byte images the reference would
 * upload to its iGPU (reference: RayAccelerator/Scene.cpp:342-346).
 */
#ifndef RACC_REF_SHIM_CL_H
#define RACC_REF_SHIM_CL_H
#include <stdlib.h>
#include <string.h>

typedef struct _cl_context* cl_context;
typedef struct _cl_program* cl_program;
typedef struct _cl_kernel* cl_kernel;
typedef struct _cl_command_queue* cl_command_queue;
typedef struct _cl_device_id* cl_device_id;
typedef int cl_int;
typedef unsigned cl_uint;
typedef unsigned long cl_mem_flags;

struct _cl_mem {
	void* data;
	size_t size;
};
typedef struct _cl_mem* cl_mem;

#define CL_MEM_READ_ONLY (1 << 2)
#define CL_MEM_USE_HOST_PTR (1 << 3)
#define CL_MEM_COPY_HOST_PTR (1 << 5)
#define CL_RGBA 0x10B5
#define CL_FLOAT 0x10DE
#define CL_MEM_OBJECT_IMAGE2D 0x10F1

typedef struct { unsigned image_channel_order, image_channel_data_type; } cl_image_format;
typedef struct {
	unsigned image_type;
	size_t image_width, image_height, image_depth, image_array_size, image_row_pitch, image_slice_pitch;
	unsigned num_mip_levels, num_samples;
	cl_mem buffer;
} cl_image_desc;

static inline cl_mem clCreateBuffer(cl_context, cl_mem_flags, size_t size, void* host, cl_int*) {
	cl_mem m = (cl_mem)malloc(sizeof(struct _cl_mem));
	m->data = malloc(size ? size : 1);
	m->size = size;
	if (host && size)
		memcpy(m->data, host, size);
	return m;
}

static inline cl_mem clCreateImage(cl_context c, cl_mem_flags f, const cl_image_format*, const cl_image_desc* d, void* host, cl_int* e) {
	return clCreateBuffer(c, f, d->image_row_pitch * d->image_height, host, e);
}

static inline cl_int clReleaseMemObject(cl_mem m) {
	free(m->data);
	free(m);
	return 0;
}
#endif
