/*
 * The reference's integrator.hpp drags in gpu_wrappers/cl_context.hpp -> CL/cl.hpp,
 * whose inline RAII wrappers reference two OpenCL entry points.  No OpenCL object
 * is ever created by the CPU harness, so these are never called; they only satisfy
 * the dynamic linker (there is no OpenCL platform in this image).
 * TEST INFRASTRUCTURE.
 */
int clReleaseCommandQueue(void* q) { (void)q; return 0; }
int clReleaseContext(void* c) { (void)c; return 0; }
