// does not compile: the error must be reported, not crash
extern "C" __global__ void linearKernel(float* out) { out[0] = undefined_symbol; }
