"""numpy views of the reference's buffer layouts (the byte formats on both sides of the path).

  NODE      LinearBVHNode  32 B  include/lens_trace/acceleration_structure_explicit.h:20-32
  PRIM      Primitive      76 B  include/lens_trace/acceleration_structure_explicit.h:34-42
  MATERIAL  Material       32 B  include/lens_trace/model.h:26-31
  LIGHTS    LightContainer 260 B include/lens_trace/acceleration_structure_explicit.h:44-47
  CAMERA    camera buffer  28 B  src/camera.cpp:14-19
"""
import numpy as np

NODE = np.dtype([("min", "<f4", 3), ("max", "<f4", 3), ("offset", "<i4"), ("count", "<u2"), ("axis", "u1"),
                 ("pad", "u1")])
PRIM = np.dtype([("a", "<f4", 3), ("b", "<f4", 3), ("c", "<f4", 3), ("na", "<f4", 3), ("nb", "<f4", 3),
                 ("nc", "<f4", 3), ("mat", "<i4")])
MATERIAL = np.dtype([("diffuse", "<f4", 3), ("ior", "<f4"), ("dissolve", "<f4"), ("emission", "<f4", 3)])
LIGHTS = np.dtype([("count", "<u4"), ("prims", "<u4", 64)])
CAMERA = np.dtype([("pos", "<f4", 3), ("yaw", "<f4"), ("pitch", "<f4"), ("roll", "<f4"), ("frameCount", "<u4")])

assert NODE.itemsize == 32 and PRIM.itemsize == 76 and MATERIAL.itemsize == 32
assert LIGHTS.itemsize == 260 and CAMERA.itemsize == 28

KERNEL_BASIC_CU = 0
KERNEL_BASIC_CL = 1
KERNEL_CUSTOM_BARY = 2
KERNEL_LIGHTING25 = 3
KERNEL_ACCUMULATOR = 4
KERNEL_GI25 = 5
KERNEL_GI = 6
KERNEL_COUNT = 7

ACCUM_NONE = 0
ACCUM_RUNNING_MEAN = 1
ACCUM_WEIGHTED_SUM = 2

SPLIT_AUTO = 0
SPLIT_SAMPLES = 1
SPLIT_TILES = 2

FLAG_STATS = 1
FLAG_CULL = 2
FLAG_MEGAKERNEL = 4
FLAG_WAVEFRONT = 8
FLAG_SERIAL = 32  # wavefront: no overlap of consecutive batches (timing kernels in isolation; same output)
FLAG_STATS_TRACED = 128  # with FLAG_STATS: count the rays actually traced (light-hit retraces folded)
FLAG_NO_STREAM = 64  # large scenes, deterministic kernels: one thread per pixel instead of tile-fetching persistent warps (same output)
FLAG_NO_THREADED = 16  # small scenes: stack kernels instead of the stackless threaded tree (same output)


def make_camera(x, y, z, yaw=0.0, frame_count=0):
    cam = np.zeros(1, dtype=CAMERA)
    cam["pos"][0] = (x, y, z)
    cam["yaw"] = yaw
    cam["frameCount"] = frame_count
    return cam


class SceneBuffers:
    """The five flat buffers a kernel launch reads, as contiguous numpy arrays (owned copies)."""

    def __init__(self, nodes, prims, materials, lights):
        self.nodes = np.ascontiguousarray(nodes, dtype=NODE)
        self.prims = np.ascontiguousarray(prims, dtype=PRIM)
        self.materials = np.ascontiguousarray(materials, dtype=MATERIAL)
        self.lights = np.ascontiguousarray(lights, dtype=LIGHTS).reshape(1)

    def save(self, path):
        np.savez_compressed(path, nodes=self.nodes, prims=self.prims, materials=self.materials, lights=self.lights)

    @staticmethod
    def load(path):
        z = np.load(path)
        return SceneBuffers(z["nodes"], z["prims"], z["materials"], z["lights"])
