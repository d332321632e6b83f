// lt-pipeline: basic_cl
// Kernel descriptor for the B200 renderer.  The reference compiles the file named by
// RenderProperties*::kernelFilePath at run time; this renderer maps the same path to a built-in,
// hand-written sm_100a pipeline (lens_trace_b200/csrc/lt_kernels.cu) chosen by the file's base name
// (and, for global_illumination.cl, by this tag / a SAMPLE_COUNT definition).
// Pipeline: primary rays, lens refraction, flat diffuse shade
