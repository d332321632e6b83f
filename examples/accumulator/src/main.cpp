// accumulator example (headless): direct lighting, one sample per frame, running mean.
//   usage: accumulator [--size W H] [--frames N] [--per-frame] [--out file.ppm|.pfm]   (from the repository root)
#include "../../progressive_common.h"

int main(int argc, char** argv) {
  return runProgressiveExample("examples/accumulator/resources/kernels/accumulator.cl", argc, argv);
}
