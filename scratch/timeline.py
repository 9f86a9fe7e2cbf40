"""Kernel timeline of CUDA-graph replays of the bench step (torch.profiler / CUPTI): name, stream, start, duration.
    python scratch/timeline.py [SAGE|GAT] [hidden] [layers]  ->  gpurun_out/timeline_<...>.csv + a text summary"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402


def main():
    backbone = sys.argv[1] if len(sys.argv) > 1 else "SAGE"
    hidden = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    layers = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    sys.argv = [sys.argv[0]]
    args = bench.parse()
    world, rank = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0"))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    if world > 1:
        import datetime
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=120))
        args.no_parity = True
    r = bench.Runner(args, dev, rank, world, "single" if world == 1 else "strong", backbone=backbone, hidden=hidden, layers=layers)
    r.prepare()
    ms, _ = r.time_resident(10)
    print(f"{backbone} h={hidden} L={layers} world={world}: {ms:.3f} ms/step ({r.graph_note})")
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        for _ in range(3):
            r.run(r.x_dev)
        torch.cuda.synchronize()
    evs = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
    evs.sort(key=lambda e: e.time_range.start)
    if not evs:
        print("no CUDA events captured")
        return
    # keep the last replay: find the largest gap-free window at the end
    t_end = max(e.time_range.end for e in evs)
    step_us = ms * 1e3
    last = [e for e in evs if e.time_range.start >= t_end - step_us * 1.02]
    t0 = min(e.time_range.start for e in last)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    if rank != 0:
        r.close()
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
        return
    path = os.path.join(ROOT, "gpurun_out", f"timeline_{backbone}_h{hidden}_L{layers}" + (f"_n{world}" if world > 1 else "") + ".csv")
    streams = {}
    with open(path, "w") as f:
        f.write("start_us,dur_us,stream,name\n")
        for e in last:
            st = streams.setdefault(getattr(e, "device_resource_id", getattr(e, "stream", 0)), len(streams))
            f.write(f"{e.time_range.start - t0:.1f},{e.time_range.end - e.time_range.start:.1f},{st},\"{e.name[:90]}\"\n")
    print("wrote", path, len(last), "kernels; streams:", len(streams))
    # exposure summary: time covered by 'big' kernels (>= 100 us) vs the rest
    big = [(e.time_range.start - t0, e.time_range.end - t0) for e in last if e.time_range.end - e.time_range.start >= 100]
    big.sort()
    covered, cur_s, cur_e = 0.0, None, None
    for s, e in big:
        if cur_e is None or s > cur_e:
            if cur_e is not None:
                covered += cur_e - cur_s
            cur_s, cur_e = s, e
        else:
            cur_e = max(cur_e, e)
    if cur_e is not None:
        covered += cur_e - cur_s
    span = max(e.time_range.end for e in last) - t0
    print(f"span {span:.0f} us; union of kernels >= 100 us: {covered:.0f} us; not covered by a big kernel: {span - covered:.0f} us")
    gaps = []
    prev_end = 0.0
    for s, e in big:
        if s - prev_end > 15:
            gaps.append((prev_end, s))
        prev_end = max(prev_end, e)
    for a, b in gaps:
        names = [f"{x.name[:40]}({x.time_range.end - x.time_range.start:.0f})" for x in last
                 if x.time_range.start - t0 < b and x.time_range.end - t0 > a and x.time_range.end - x.time_range.start < 100]
        print(f"  gap {a:.0f}-{b:.0f} us ({b - a:.0f}): " + ", ".join(names[:14]))
    if world > 1:
        r.close()
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
