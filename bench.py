#!/usr/bin/env python3
"""bench.py -- the measurement contract of this repo.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl reference]

Default workload = BASELINE.json configs[1]: Cornell box, the global_illumination example kernel
(one sample per frame, seed = frameCount), 1920x1080, 64 frames kept as a running mean on the device,
4 bounces, one B200.  A "step" is one such 64-spp frame.  metric = Mrays/s (every BVH traversal call is
one ray: primary + shadow + GI extension), ms_per_step = ms/frame 1080p 64spp.

N > 1 (torchrun, one rank per GPU): sample split -- rank r renders frames r, r+N, ... of a 64*N-spp
frame into its own FP32 accumulator with weight 1/(64N); one NCCL all-reduce(sum) per step combines
them (weak scaling: per-GPU work is fixed).  Time = CUDA events on the stream the kernels run on, max
over ranks.

--impl reference: the reference has no CPU implementation of this path and its OpenCL kernels cannot
run here (no OpenCL ICD), so this arm times the CPU restatement of the same kernel (oracle/, "port")
on all host cores, on a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from lens_trace_b200 import layouts as L  # noqa: E402

WORKLOADS = {
    # name: (model, kernel, width, height, frames, max_ray_depth, description)
    "cornell_gi_1080p_64spp": ("cornell_box", L.KERNEL_GI, 1920, 1080, 64, 4,
                               "Cornell box, global_illumination example kernel, 1080p, 64 spp running mean, 4 bounces"),
    "cornell_primary_512": ("cornell_box", L.KERNEL_BASIC_CL, 512, 512, 1, 0,
                            "Cornell box, basic kernel, 512x512 primary rays 1 spp"),
    "cornell_custom_1080p_16": ("cornell_box", L.KERNEL_CUSTOM_BARY, 1920, 1080, 16, 0,
                                "custom_kernel (barycentric) on the Cornell OBJ, 1080p, 16 accumulated frames"),
    "synth1m_shadow_1080p": ("synth:707", L.KERNEL_ACCUMULATOR, 1920, 1080, 1, 0,
                             "synthetic 1M-triangle mesh, primary + shadow rays, 1080p 1 spp"),
    "synth1m_gi_1080p_16spp": ("synth:707", L.KERNEL_GI, 1920, 1080, 16, 4,
                               "synthetic 1M-triangle mesh, GI 4 bounces, 1080p 16 spp"),
    "synth5m_gi_4k_256spp": ("synth:1581", L.KERNEL_GI, 3840, 2160, 256, 4,
                             "synthetic 5M-triangle mesh, GI 4 bounces, 4K 256 spp"),
}
DEFAULT_WORKLOAD = "cornell_gi_1080p_64spp"


def load_scene(model):
    from lens_trace_b200 import host
    if model.startswith("synth:"):
        n = int(model.split(":")[1])
        path = "/tmp/lt_synth_%d_%d.obj" % (n, os.getpid())
        host.write_synthetic_scene(path, n, 0x5EED)
        sb = host.load_scene_buffers(path)
        for p in (path, path[:-4] + ".mtl"):
            try:
                os.remove(p)
            except OSError:
                pass
        return sb
    return host.load_scene_buffers(os.path.join(ROOT, "resources", "models", model + ".obj"))


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return {"hbm_gbs": d.get("hbm_gbs", 6650.0), "sm_max_mhz": d.get("sm_max_mhz", 1965.0), "source": "measured"}
    return {"hbm_gbs": 6650.0, "sm_max_mhz": 1965.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe): one
    nvidia-smi process streaming a sample every 50 ms between start() and summary()."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc = index, None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            time.sleep(0.15)  # let the first sample land before the timed region starts
        except Exception:
            self.proc = None

    def summary(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.06)
        self.proc.terminate()
        try:
            out = self.proc.communicate(timeout=5)[0]
        except Exception:
            out = ""
        samples = []
        for line in out.splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) >= 6:
                try:
                    float(f[0])
                    samples.append(f)
                except ValueError:
                    pass
        if not samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi gave no samples"]}
        sm = sorted(float(s[0]) for s in samples)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in samples)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_min_mhz": sm[0], "sm_max_mhz": float(samples[0][1]), "reasons": reasons,
                "samples": len(sm)}


def cpu_oracle_rate(sb, kernel, w, h, depth, frame_list, threads=0):
    """Mrays/s of the CPU restatement on `frame_list` full frames; returns (mrays_s, rays, seconds, cores)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lt_oracle as O
    cores = os.cpu_count() or 1
    rays, t = 0, 0.0
    for f in frame_list:
        cam = L.make_camera(0, 2.5, -50, 0.0, f)
        t0 = time.perf_counter()
        _, st = O.render(kernel, sb, cam, w, h, max_ray_depth=depth if depth else 16, threads=threads, with_stats=True)
        t += time.perf_counter() - t0
        rays += st.rays
    return rays / t / 1e6, rays, t, (threads if threads > 0 else cores)


def cpu_reference_text_seconds(sb, kernel, w, h, depth, frame_list):
    """Seconds the reference's OWN kernel file takes on the host cores for `frame_list` full frames: its OpenCL C
    text compiled for the CPU through oracle/cl_shim (oracle/_ref/libltref_cl.so, built where /root/reference
    exists and shipped with the repo snapshot) -- the stand-in for the north star's "OpenCL backend on a CPU
    device (POCL)", which cannot be installed here.  None when the library is absent or the kernel has no .cl file."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import lt_ref_cl as R
    if not R.available() or not (1 <= kernel <= 6):
        return None
    t = 0.0
    for f in frame_list:
        cam = L.make_camera(0, 2.5, -50, 0.0, f)
        t0 = time.perf_counter()
        R.render(kernel, sb, cam, w, h, max_ray_depth=depth if depth else 16, threads=0)
        t += time.perf_counter() - t0
    return t


def run_reference(args, rank, world):
    """--impl reference: the reference's kernel on the CPU, all host cores, bounded sample per step.  The reference's
    own kernel text (oracle/_ref/libltref_cl.so) when it is present, else the CPU port (oracle/lt_oracle.c); the ray
    count of a frame always comes from the port's counters."""
    if rank != 0:
        return
    model, kernel, w, h, frames, depth, desc = WORKLOADS[args.workload]
    sb = load_scene(model)
    per_step = max(1, min(frames, 2))
    cpu_oracle_rate(sb, kernel, w, h, depth, [0])  # warm-up (page in, build)
    use_text = cpu_reference_text_seconds(sb, kernel, w, h, depth, [0]) is not None
    for i in range(max(0, args.warmup - 1)):
        cpu_oracle_rate(sb, kernel, w, h, depth, [i])
    rays, secs, port_secs, cores = 0, 0.0, 0.0, 1
    for s in range(args.steps):
        fl = [s * per_step + k for k in range(per_step)]
        _, r, t, cores = cpu_oracle_rate(sb, kernel, w, h, depth, fl)
        rays += r
        port_secs += t
        secs += cpu_reference_text_seconds(sb, kernel, w, h, depth, fl) if use_text else t
    value = rays / secs / 1e6
    sample = "%d of the %d frames per step at full %dx%d (frameCount = step*%d + k)" % (per_step, frames, w, h, per_step)
    line = {
        "impl": "reference", "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": secs / args.steps * 1e3 * (frames / per_step),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": args.workload, "description": desc, "width": w, "height": h, "frames_per_step": frames,
                   "max_ray_depth": depth, "note": "ms_per_step extrapolated from the bounded sample to the whole step"},
        "cpu_baseline": {"value": value, "unit": "Mrays/s", "cores": cores, "kind": "reference" if use_text else "port",
                         "sample": sample, "port_value": rays / port_secs / 1e6,
                         "what": "the reference's own kernel file compiled for the CPU (oracle/cl_shim), std::thread over "
                                 "image rows; port_value = the C restatement (oracle/lt_oracle.c)" if use_text else
                                 "C restatement of the reference kernel (oracle/lt_oracle.c), pthreads over image rows"},
        "e2e": {"value": value, "unit": "Mrays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def reference_cuda_kernel_rate(sb, w, h):
    """The reference's own CUDA backend on this GPU, primary rays (its only CUDA kernel, basic.cu), same scene
    and size -- context for the north star's 10x target.  Two numbers, as SURVEY.md 8(d) asks:
      kernel only : NVRTC build of its basic.cu launched on resident buffers, CUDA events (ltref_time_kernel)
      as shipped  : wall time of RendererCUDA::render() -- per call it JIT-loads the module, allocates and
                    uploads five buffers, launches, synchronises and copies the frame back (pageable memory)
    None when oracle/_ref is absent."""
    import ctypes as C
    lib_path = os.path.join(ROOT, "oracle", "_ref", "libltref.so")
    kpath = os.path.join(ROOT, "oracle", "_ref", "resources", "kernels", "cuda", "basic.cu")
    if not (os.path.exists(lib_path) and os.path.exists(kpath)):
        return None
    try:
        lib = C.CDLL(lib_path)
        vp, u64 = C.c_void_p, C.c_uint64
        lib.ltref_renderer_cuda_create.restype = vp
        lib.ltref_time_kernel.restype = C.c_double
        lib.ltref_time_kernel.argtypes = [C.c_char_p, C.c_char_p, vp, u64, vp, u64, vp, u64, vp, u64, vp, u64, u64, u64,
                                          C.c_uint, C.c_uint, C.c_int, C.c_int, vp]
        renderer = lib.ltref_renderer_cuda_create()
        cam = L.make_camera(0, 2.5, -50)
        res = {}
        for bx, by in ((32, 1), (8, 8)):
            ms = lib.ltref_time_kernel(kpath.encode(), b"linearKernel", sb.nodes.ctypes.data, sb.nodes.nbytes,
                                       sb.prims.ctypes.data, sb.prims.nbytes, sb.materials.ctypes.data,
                                       sb.materials.nbytes, sb.lights.ctypes.data, sb.lights.nbytes, cam.ctypes.data,
                                       w, h, 3, bx, by, 3, 10, None)
            if ms > 0:
                res["block_%dx%d" % (bx, by)] = {"ms": ms, "mrays_s": w * h / ms / 1e3}
        # as shipped: the reference's own objects (its Model/AS of the Cornell box) through RendererCUDA::render
        obj = os.path.join(ROOT, "oracle", "_ref", "resources", "models", "cornell_box.obj")
        if os.path.exists(obj) and len(sb.prims) == 42:
            for n in ("ltref_model_create", "ltref_as_create", "ltref_camera_create"):
                getattr(lib, n).restype = vp
            lib.ltref_model_create.argtypes = [C.c_char_p]
            lib.ltref_as_create.argtypes = [vp]
            lib.ltref_camera_create.argtypes = [C.c_float] * 4
            lib.ltref_render_cuda.argtypes = [vp, C.c_char_p, C.c_int, C.c_int, u64, u64, u64, u64, u64, vp, u64, vp, vp,
                                              vp]
            lib.ltref_render_cuda.restype = None
            model = lib.ltref_model_create(obj.encode())
            accel = lib.ltref_as_create(model)
            rcam = lib.ltref_camera_create(0, 2.5, -50, 0)
            out = np.zeros((h, w, 3), np.float32)
            times = []
            for i in range(6):
                t0 = time.perf_counter()
                lib.ltref_render_cuda(renderer, kpath.encode(), 0, 0, 0, 0, w, h, 3, out.ctypes.data, out.nbytes, accel,
                                      model, rcam)
                times.append((time.perf_counter() - t0) * 1e3)
            res["render_as_shipped_ms"] = sorted(times[1:])[len(times[1:]) // 2]
        return res or None
    except Exception as e:  # pragma: no cover
        return {"error": str(e)}


def this_repo_render_call_ms(w, h):
    """Wall time of this repo's RendererCUDA::render() for the same call (scene cached after the first call,
    kernel, D2H into a malloc'ed host buffer)."""
    from lens_trace_b200 import host
    cwd = os.getcwd()
    os.chdir(ROOT)
    try:
        cam = host.Camera(0, 2.5, -50, 0)
        model = host.Model("resources/models/cornell_box.obj")
        accel = host.AccelerationStructure(model)
        r = host.Renderer(host.PLATFORM_CUDA)
        out = np.zeros((h, w, 3), np.float32)
        times = []
        for i in range(6):
            t0 = time.perf_counter()
            r.render("resources/kernels/cuda/basic.cu", w, h, accel, model, cam, out=out)
            times.append((time.perf_counter() - t0) * 1e3)
        r.close(); accel.close(); model.close(); cam.close()
        return sorted(times[1:])[len(times[1:]) // 2]
    finally:
        os.chdir(cwd)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="weak: every GPU renders the workload's frames (N x the samples); strong: the workload's "
                         "frames are divided among the GPUs (BASELINE configs[4]: 256 spp split over 1/2/4/8)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-ref-cuda", action="store_true")
    ap.add_argument("--no-cull", action="store_true", help="skip the secondary opt-in culled measurement")
    ap.add_argument("--no-lbvh", action="store_true", help="skip the secondary opt-in GPU-built LBVH measurement")
    args = ap.parse_args()

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch
    import torch.distributed as dist
    from lens_trace_b200 import capi

    if not torch.cuda.is_available():
        sys.exit("bench.py: no CUDA device; the B200 path has no CPU fallback (use --impl reference for the CPU port)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # keep stdout to the one JSON line: NCCL writes its version banner (and anything else it logs) to stdout
        # unless it is given a file, and it honours NCCL_DEBUG_FILE only above the VERSION level
        if os.environ.get("NCCL_DEBUG", "").upper() in ("", "VERSION"):
            os.environ["NCCL_DEBUG"] = "WARN"
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=dev)

    model, kernel, w, h, frames, depth, desc = WORKLOADS[args.workload]
    if args.scaling == "strong":
        if frames % world:
            sys.exit("bench.py: --scaling strong needs the workload's %d frames to divide by %d GPUs" % (frames, world))
        frames //= world  # this rank's share; frame k of rank r has frameCount r + k*world, weight 1/(frames*world)
    sb = load_scene(model)
    ctx = capi.Context(local_rank)
    scene = ctx.upload(sb)
    upload_ms = ctx.stats().upload_ms
    # a non-default torch stream: lt_ctx_set_stream(NULL) would mean "the context's own stream", and
    # torch.cuda.Event only sees the stream it is recorded on
    stream = torch.cuda.Stream(device=dev)
    torch.cuda.set_stream(stream)
    ctx.set_stream(stream.cuda_stream)

    acc = torch.zeros((h, w, 3), dtype=torch.float32, device=dev)
    flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)  # > 126 MB L2
    pinned = torch.empty((h, w, 3), dtype=torch.float32).pin_memory()

    def make_step_params(flags=0):
        if world == 1:
            return capi.make_params(kernel, w, h, max_ray_depth=depth, frames=frames,
                                    accum_mode=L.ACCUM_RUNNING_MEAN, flags=flags)
        return capi.make_params(kernel, w, h, max_ray_depth=depth, frames=frames, frame_stride=world,
                                accum_mode=L.ACCUM_WEIGHTED_SUM, accum_weight=1.0 / (frames * world), flags=flags)

    cam = L.make_camera(0, 2.5, -50, 0.0, rank if world > 1 else 0)

    def step():
        """one frame of the workload: all samples of this rank, then the exchange step"""
        if world > 1:
            acc.zero_()
        ctx.render_device(scene, cam, make_step_params(), acc.data_ptr(), sync=False)
        if world > 1:
            dist.all_reduce(acc)

    # counters of one step (untimed): rays and the reference-order node/triangle tests
    ctx.render_device(scene, cam, make_step_params(L.FLAG_STATS), acc.data_ptr(), sync=True)
    st = ctx.stats()
    counts = torch.tensor([st.rays, st.node_tests, st.tri_tests], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(counts)
    rays, node_tests, tri_tests = [float(x) for x in counts.tolist()]
    # The camera ray of a pixel is the same in every frame (the reference has no sub-pixel jitter), so the wavefront
    # pipeline traces it once per pixel per step and starts all `frames` paths of the pixel from that hit record.
    # rays / node_tests / tri_tests above are the reference's counts (one primary ray per pixel per frame); the
    # traversal kernels actually execute (frames - 1) primary rays per pixel fewer.  Counted with the primary-only
    # kernel (basic.cl: one camera ray per pixel; the bench scenes have no lens material).
    ctx.render_device(scene, cam, capi.make_params(L.KERNEL_BASIC_CL, w, h, flags=L.FLAG_STATS), acc.data_ptr(), sync=True)
    sp = ctx.stats()
    primary_counts = (float(sp.rays), float(sp.node_tests), float(sp.tri_tests))

    for _ in range(max(3, args.warmup)):
        step()
    torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    events = []
    kernel_events = []
    trace_ms_total, trace_launches = 0.0, 0
    for _ in range(args.steps):
        flush.fill_(1.0)  # L2 flush between timed iterations (untimed)
        e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
        e0.record(stream)
        if world > 1:
            acc.zero_()
        ctx.render_device(scene, cam, make_step_params(), acc.data_ptr(), sync=False)
        e1.record(stream)
        if world > 1:
            dist.all_reduce(acc)
        e2.record(stream)
        events.append((e0, e2))
        kernel_events.append((e0, e1))
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    clocks = sampler.summary()
    launches_per_step = ctx.stats().kernel_launches
    # The dominant kernels timed in isolation, live, for the roofline: in the timed steps above consecutive batches
    # of the wavefront pipeline overlap on two streams (a trace kernel beside another batch's shade kernel), so a
    # launch's own duration is only defined with the overlap off (LT_FLAG_SERIAL: same kernels, same output, one
    # stream).  Synchronous call: the library brackets every traversal launch with CUDA events on this stream and
    # sums them (lt_stats.trace_ms).
    iso_steps = max(1, min(args.steps, 3))
    iso_pipeline_ms = 0.0
    for _ in range(iso_steps):
        flush.fill_(1.0)
        if world > 1:
            acc.zero_()
        ctx.render_device(scene, cam, make_step_params(L.FLAG_SERIAL), acc.data_ptr(), sync=True)
        st_step = ctx.stats()
        trace_ms_total += st_step.trace_ms
        trace_launches += st_step.trace_launches
        iso_pipeline_ms += st_step.kernel_ms
    if world > 1:
        acc.zero_()
        ctx.render_device(scene, cam, make_step_params(), acc.data_ptr(), sync=True)  # leave acc as a step leaves it
        dist.all_reduce(acc)
    total_ms = sum(a.elapsed_time(b) for a, b in events)
    kernel_ms = sum(a.elapsed_time(b) for a, b in kernel_events)
    t = torch.tensor([total_ms, kernel_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    total_ms, kernel_ms = t.tolist()
    ms_per_step = total_ms / args.steps
    value = rays / (ms_per_step * 1e-3) / 1e6

    # end to end through the public C-ABI with host buffers (lt_render: H2D of the camera, kernels, D2H)
    e2e_times = []
    for i in range(max(2, min(args.steps, 3)) + 1):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if world == 1:
            ctx.render_into(scene, cam, make_step_params(), pinned.data_ptr())
        else:
            acc.zero_()
            ctx.render_device(scene, cam, make_step_params(), acc.data_ptr(), sync=False)
            dist.all_reduce(acc)
            pinned.copy_(acc, non_blocking=False)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        if i > 0:
            e2e_times.append(dt)
    e2e_t = torch.tensor([sum(e2e_times) / len(e2e_times)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(e2e_t, op=dist.ReduceOp.MAX)
    e2e_s = e2e_t.item()

    if rank == 0:
        pk = peaks()
        sm_count = st.sm_count or 148
        fp32_peak = sm_count * 128 * pk["sm_max_mhz"] * 1e6  # lane-instructions / s
        # roofline of ONE GPU's kernels: the counters were summed over ranks
        shared_primary = launches_per_step > 1 and frames > 1  # the wavefront pipeline (exact mode) does this
        skipped = (frames - 1) * world if shared_primary else 0
        rays_traced = rays - skipped * primary_counts[0]
        node_traced = node_tests - skipped * primary_counts[1]
        tri_traced = tri_tests - skipped * primary_counts[2]
        if kernel <= 2 and frames > 1:
            # deterministic kernels: every frame is the same image, k_flat traces once and applies the frame
            # combiner `frames` times in registers
            shared_primary = True
            rays_traced, node_traced, tri_traced = rays / frames, node_tests / frames, tri_tests / frames
        # algorithmic work of the rays the timed traversal launches actually trace (32 B and 12
        # lane-instructions per box test, 36 B and 45 per triangle test, counted in the reference's traversal order)
        alg_instr = (12.0 * node_traced + 45.0 * tri_traced) / world
        alg_bytes = (32.0 * node_traced + 36.0 * tri_traced) / world
        ref_alg_bytes = (32.0 * node_tests + 36.0 * tri_tests) / world
        # the dominant kernels are the traversal kernels; their summed duration per step, measured live
        trace_s = trace_ms_total / iso_steps * 1e-3
        k_s = trace_s if trace_s > 0 else kernel_ms / args.steps * 1e-3
        scene_bytes = sb.nodes.nbytes + sb.prims.nbytes
        # what the traversal kernels read: 64-byte child-pair nodes (32 B per reference node) + 48-byte triangles
        traversal_bytes = len(sb.nodes) * 32 + len(sb.prims) * 48
        # Level that bounds the node fetches.  A traversal set that fits the 126 MB L2 is served by L2/L1: ncu on the
        # 1 M-triangle mesh (112 MB, profiles/r1k_*) shows DRAM at 4 % and the L1 data pipe 80 % busy, like the 21 KB
        # Cornell tree (84 %) -- each lane of a warp gathers its own 32-byte record.  Such workloads are rated
        # against the nominal L1 rate; only sets beyond L2 are rated against the measured HBM copy bandwidth.
        level = "hbm" if traversal_bytes > 126e6 else "l1"
        l1_peak = sm_count * 128 * pk["sm_max_mhz"] * 1e6 / 1e9  # GB/s, nominal 128 B/clk/SM
        bw_peak = {"hbm": pk["hbm_gbs"], "l1": l1_peak}[level]
        fp32 = {"achieved": alg_instr / k_s / 1e12, "peak": fp32_peak / 1e12, "unit": "T lane-instr/s",
                "frac": alg_instr / k_s / fp32_peak}
        fetch = {"level": level, "achieved": alg_bytes / k_s / 1e9, "peak": bw_peak, "unit": "GB/s",
                 "frac": (alg_bytes / k_s / 1e9 / bw_peak) if bw_peak else None,
                 "peak_source": ("MEASURED_PEAKS.json (%s)" % pk["source"]) if level == "hbm" else
                 "nominal 128 B/clk/SM x SMs x sm_max_mhz (traversal set of %d bytes is L1/L2-resident)" % traversal_bytes}
        t_fp32 = alg_instr / fp32_peak
        t_fetch = (alg_bytes / (bw_peak * 1e9)) if bw_peak else 0.0
        if t_fetch >= t_fp32:
            roof = {"bound": level, "achieved": fetch["achieved"], "peak": fetch["peak"], "unit": "GB/s",
                    "frac": fetch["frac"], "traffic": None}
        else:
            roof = {"bound": "fp32", "achieved": fp32["achieved"], "peak": fp32["peak"], "unit": "T lane-instr/s",
                    "frac": fp32["frac"], "traffic": None}
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            tr = json.load(open(tpath)).get(args.workload)
            if tr:
                roof["traffic"] = tr["bytes"]
                roof["traffic_source"] = tr["capture"]
        pipeline = "wavefront (k_wf_primary_trace once, then per batch k_wf_primary + (k_wf_trace, k_wf_shade) x rounds + k_wf_accumulate; two batches in flight on two streams)" \
            if launches_per_step > 1 else ("k_path" if kernel >= 3 else "k_flat")
        roof.update({"kernel": "k_wf_primary_trace + k_wf_trace (traversal kernels of the wavefront pipeline)"
                     if launches_per_step > 1 else pipeline,
                     "pipeline": pipeline,
                     "timing": "traversal launches timed live with CUDA events, kernels in isolation (LT_FLAG_SERIAL: "
                               "one stream, %d steps); the timed steps overlap consecutive batches on two streams" % iso_steps,
                     "traversal_ms_per_step": trace_ms_total / iso_steps,
                     "traversal_launches_per_step": trace_launches / iso_steps,
                     "traversal_avg_launch_ms": trace_ms_total / max(1, trace_launches),
                     "traversal_share_of_step": trace_ms_total / max(1e-9, iso_pipeline_ms),
                     "pipeline_ms_per_step_in_isolation": iso_pipeline_ms / iso_steps,
                     "pipeline_ms_per_step": kernel_ms / args.steps,
                     "kernels_per_step": launches_per_step,
                     "algorithmic_units": "per step (all traversal launches of one step), per GPU; rays actually traced",
                     "algorithmic_bytes_per_step": alg_bytes, "algorithmic_fp32_instr_per_step": alg_instr,
                     "rays_traced_per_step": rays_traced / world,
                     "reference_rays_per_step": rays / world,
                     "reference_algorithmic_bytes_per_step": ref_alg_bytes,
                     "primary_rays": "traced once per pixel per step and shared by the step's %d frames "
                                     "(identical camera ray in every frame); value counts the reference's rays" % frames
                     if shared_primary else "one per pixel per frame",
                     "node_tests_per_ray": node_tests / rays, "tri_tests_per_ray": tri_tests / rays,
                     "fp32_issue": fp32, "node_fetch": fetch,
                     "hbm_view": {"achieved": alg_bytes / k_s / 1e9, "peak": pk["hbm_gbs"], "unit": "GB/s",
                                  "frac": alg_bytes / k_s / 1e9 / pk["hbm_gbs"], "peak_source": pk["source"]}})
        line = {
            "metric": "Mrays/s", "value": value, "unit": "Mrays/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(3, args.warmup), "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": args.scaling,
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": args.workload, "description": desc, "width": w, "height": h,
                       "frames_per_step_per_gpu": frames, "max_ray_depth": depth, "rays_per_step": rays,
                       "parallelism": "spp-split x%d + 1 all-reduce/step" % world if world > 1 else "single GPU",
                       "l2": "flushed between timed steps (256 MB write); scene is %d bytes" % scene_bytes,
                       "scene_upload_ms": upload_ms},
            "roofline": roof,
            "e2e": {"value": rays / e2e_s / 1e6, "unit": "Mrays/s", "ms_per_step": e2e_s * 1e3,
                    "h2d_bytes_per_step": 28, "d2h_bytes_per_step": int(w * h * 3 * 4),
                    "api": "lt_render (C-ABI, host output buffer)" if world == 1 else
                    "lt_render_device + NCCL all-reduce + D2H"},
            "gpu_launches": args.steps * launches_per_step,
            "clocks": clocks,
        }
        if world == 1 and not args.no_cpu_baseline:
            # bounded sample: about 10-30 s of CPU work on the box's cores (one frame first to size it)
            _, _, t1, _ = cpu_oracle_rate(sb, kernel, w, h, depth, [0])
            per = max(1, min(frames, int(15.0 / max(t1, 1e-3))))
            v, r, secs, cores = cpu_oracle_rate(sb, kernel, w, h, depth, list(range(per)))
            line["cpu_baseline"] = {"value": v, "unit": "Mrays/s", "cores": cores, "kind": "port",
                                    "sample": "frames 0..%d of the %d at full %dx%d (%.1f s of CPU work)" % (
                                        per - 1, frames, w, h, secs)}
            # the reference's own kernel text on the same cores, on a smaller sample of the same frames
            per_text = max(1, per // 2)
            tsecs = cpu_reference_text_seconds(sb, kernel, w, h, depth, list(range(per_text)))
            if tsecs is not None:
                _, r_text, _, _ = cpu_oracle_rate(sb, kernel, w, h, depth, list(range(per_text))) if per_text != per else (0, r, 0, 0)
                line["cpu_baseline"] = {
                    "value": r_text / tsecs / 1e6, "unit": "Mrays/s", "cores": cores, "kind": "reference",
                    "sample": "frames 0..%d of the %d at full %dx%d (%.1f s of CPU work)" % (per_text - 1, frames, w, h, tsecs),
                    "what": "the reference's own kernel file (OpenCL C) compiled for the CPU through oracle/cl_shim, "
                            "std::thread over image rows -- stand-in for the OpenCL backend on a POCL CPU device",
                    "port_value": v, "port_sample": "frames 0..%d (%.1f s), C restatement oracle/lt_oracle.c" % (per - 1, secs)}
        if world == 1 and not args.no_cull:
            # OPT-IN culled traversal (LT_FLAG_CULL): reported beside the headline, never instead of it.
            # It does less work than the reference's traversal; identity of the full-size output is checked here.
            ctx.set_stream(stream.cuda_stream)
            exact = acc.clone()
            pc = make_step_params(L.FLAG_CULL)
            ctx.render_device(scene, cam, pc, acc.data_ptr(), sync=True)
            identical = bool(torch.equal(acc.view(torch.int32), exact.view(torch.int32)))
            times = []
            for _ in range(max(2, min(args.steps, 3))):
                flush.fill_(1.0)
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c0.record(stream)
                ctx.render_device(scene, cam, pc, acc.data_ptr(), sync=False)
                c1.record(stream)
                torch.cuda.synchronize()
                times.append(c0.elapsed_time(c1))
            ctx.render_device(scene, cam, make_step_params(L.FLAG_CULL | L.FLAG_STATS), acc.data_ptr(), sync=True)
            sc_ = ctx.stats()
            cms = sum(times) / len(times)
            line["culled_opt_in"] = {
                "flag": "LT_FLAG_CULL (off by default)", "ms_per_step": cms, "value": rays / (cms * 1e-3) / 1e6,
                "unit": "Mrays/s (same ray count as the exact run)", "output_bit_identical_to_exact": identical,
                "node_tests_per_ray_actual": sc_.node_tests / max(1, sc_.rays),
                "tri_tests_per_ray_actual": sc_.tri_tests / max(1, sc_.rays),
                "speedup_vs_exact": ms_per_step / cms}
        if world == 1 and not args.no_lbvh and len(sb.prims) >= 2:
            # OPT-IN device-built LBVH (lt_scene_build_lbvh): a different tree than the reference builder's, so it is
            # reported beside the headline only.  Same kernels, same workload; picture compared with the headline's.
            ctx.set_stream(stream.cuda_stream)
            ctx.render_device(scene, cam, make_step_params(), acc.data_ptr(), sync=True)
            exact = acc.clone()
            t0 = time.perf_counter()
            lscene = ctx.build_lbvh(sb.prims, sb.materials)
            build_wall_ms = (time.perf_counter() - t0) * 1e3
            build_dev_ms = ctx.stats().upload_ms
            times = []
            for _ in range(max(2, min(args.steps, 3)) + 1):
                flush.fill_(1.0)
                c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                c0.record(stream)
                ctx.render_device(lscene, cam, make_step_params(), acc.data_ptr(), sync=False)
                c1.record(stream)
                torch.cuda.synchronize()
                times.append(c0.elapsed_time(c1))
            lms = sum(times[1:]) / len(times[1:])
            differing = int(((acc - exact).abs().amax(dim=-1) > 1e-4 * exact.abs().amax(dim=-1).clamp_min(1e-3)).sum().item())
            ctx.render_device(lscene, cam, make_step_params(L.FLAG_STATS), acc.data_ptr(), sync=True)
            sl = ctx.stats()
            line["lbvh_opt_in"] = {
                "what": "same workload on an LBVH built on the GPU (ACCELERATION_STRUCTURE_TYPE_LBVH_B200), not the "
                        "reference builder's tree",
                "build_ms_device": build_dev_ms, "build_ms_wall": build_wall_ms, "ms_per_step": lms,
                "value": sl.rays / (lms * 1e-3) / 1e6, "unit": "Mrays/s", "speedup_vs_reference_tree": ms_per_step / lms,
                "node_tests_per_ray": sl.node_tests / max(1, sl.rays), "tri_tests_per_ray": sl.tri_tests / max(1, sl.rays),
                "pixels_differing_beyond_1e-4_rel": differing, "pixels": w * h}
            lscene.release()
        if world == 1 and not args.no_ref_cuda:
            # the reference's CUDA kernel only exists for primary rays (basic.cu); same scene, same size
            ref = reference_cuda_kernel_rate(sb, w, h)
            if ref:
                pp = capi.make_params(L.KERNEL_BASIC_CU, w, h)
                ctx.set_stream(None)
                for _ in range(3):
                    ctx.render_device(scene, L.make_camera(0, 2.5, -50), pp, acc.data_ptr(), sync=True)
                mine_ms = []
                for _ in range(10):
                    ctx.render_device(scene, L.make_camera(0, 2.5, -50), pp, acc.data_ptr(), sync=True)
                    mine_ms.append(ctx.stats().kernel_ms)
                mine = sum(mine_ms) / len(mine_ms)
                line["reference_cuda_backend"] = {
                    "what": "primary rays (basic.cu), %dx%d, same scene, same GPU" % (w, h),
                    "reference_kernel": ref, "this_repo_kernel": {"ms": mine, "mrays_s": w * h / mine / 1e3}}
                if "render_as_shipped_ms" in ref:
                    line["reference_cuda_backend"]["this_repo_render_call_ms"] = this_repo_render_call_ms(w, h)
        print(json.dumps(line), flush=True)

    scene.release()
    ctx.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
