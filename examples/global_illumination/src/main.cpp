// global_illumination example (headless): direct light + GI bounces, one sample per frame, running mean.
//   usage: global_illumination [--size W H] [--frames N] [--depth D] [--per-frame] [--out file]   (from the repository root)
#include "../../progressive_common.h"

int main(int argc, char** argv) {
  return runProgressiveExample("examples/global_illumination/resources/kernels/global_illumination.cl", argc, argv);
}
