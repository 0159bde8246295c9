/* ref_bsdf_stubs.c -- TEST INFRASTRUCTURE.  libref_bsdf.so links the reference's renderer sources without a point sampler
 * module (oracle/ref_bsdf.c supplies pointsampler() itself); these entry points of src/pointsampler.d/*.c are referenced by
 * code paths the BSDF harness never runs and only have to exist for the dynamic linker. */
void pointsampler_print_info() {}
void pointsampler_clear() {}
void pointsampler_finalize() {}
void pointsampler_prepare_frame() {}
void pointsampler_mutate() {}
int  pointsampler_accept() { return 0; }
void pointsampler_splat() {}
