#!/usr/bin/env python
"""Benchmark of the per-step track loop (BASELINE.json: track-steps/sec, TestEm3 full EM).

A bench "step" is one complete transport of one batch of synthetic primaries: the
workload BASELINE.json's metric is quoted on (configs[1]): TestEm3 full EM (Urban MSC +
energy-loss fluctuations), 10 000 1 GeV e- primaries per GPU as 100 events x 100 primaries
merged onto one state, num_track_slots = 2^20. Physics tables are the documented stand-in
(steel absorber instead of Pb; tools/make_physics.py).

  value        track-steps/s with the primaries already staged on the device
  e2e          the same through the public C-ABI call b200_run_events with HOST primaries
               (H2D of primaries, per-iteration D2H of counters, D2H of tallies, all timed)
  roofline     the dominant kernel (along-step) measured live with CUDA events
  cpu_baseline the reference's own host Stepper (oracle/_ref) on this box's host cores

    python bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference ...      (reference CPU arm)
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

import numpy as np  # noqa: E402

# GPU arm: track_order = init_charge, the reference's own default when use_device is true
# (app/celer-sim/RunnerInputIO.json.cc:113-120); CPU arm: the reference's CPU default (none).
# Same geometry, physics tables, seed and primaries.
IMAGE = os.path.join(REPO, 'data', 'images', 'testem3-initcharge.b2img')
CONFIG = os.path.join(REPO, 'data', 'images', 'testem3.json')
NUM_EVENTS = 100
PRIMARIES_PER_EVENT = 100
ENERGY_MEV = 1000.0
NUM_TRACK_SLOTS = 1 << 20
# two steppers (CUDA streams) per GPU share the 2^20 slots: the latency-bound shower tails
# of one overlap the throughput-bound iterations of the other (measured sweep:
# profiles/README_r01.md; 1 stream 6.8e8, 2 streams 7.5e8, 3 streams 7.6e8, 4 streams 7.5e8)
NUM_STREAMS = 2
ALG_BYTES_PER_TRACK_STEP = 672  # 2 * S_live, D=1, P=4
# dram__bytes_read.sum + dram__bytes_write.sum of ONE along-step launch (charged + neutral
# kernels) at a saturated iteration (all 2^20 slots live), from `ncu --set full`
NCU_TRAFFIC_BYTES_PER_LAUNCH = 135.6e6 + 86.0e6 + 90.1e6 + 27.0e6
NCU_TRAFFIC_SOURCE = ('profiles/layout_variants_r02.txt, section "rng" (the shipped layout): '
                      'k_along_step_charged<0,0> + k_along_step_neutral<0> at one saturated '
                      'iteration of this workload (launch 900 of the per-action kernels), '
                      'dram__bytes_read.sum + dram__bytes_write.sum')


# --workload: the headline (BASELINE configs[1]), simple-CMS (configs[2]) and the CMS-scale
# stand-in (configs[3]/[4]:
# tools/make_cms_scale.py, four universe levels, 1 T field, isotropic 10 GeV e-/gamma)
WORKLOADS = {
    'testem3': dict(image=IMAGE, config=CONFIG, alg_bytes=ALG_BYTES_PER_TRACK_STEP,
                    traffic=NCU_TRAFFIC_BYTES_PER_LAUNCH, traffic_source=NCU_TRAFFIC_SOURCE,
                    events=NUM_EVENTS, per_event=PRIMARIES_PER_EVENT,
                    label='TestEm3 full EM (Urban MSC + eloss fluctuations), %d x %d 1 GeV e- '
                          'primaries per GPU'),
    'cms-scale': dict(image=os.path.join(REPO, 'data', 'images', 'cms-scale.b2img'),
                      config=os.path.join(REPO, 'data', 'images', 'cms-scale.json'),
                      alg_bytes=2 * (248 + 56 * 4 + 8 * 4),  # SURVEY.md 8(d) with D=4, P=4
                      # the four phase kernels of one charged along-step at a saturated
                      # iteration (5.8e5 charged tracks): read + written bytes
                      traffic=(168.4 + 21.6 + 203.9 + 149.8 + 206.4 + 83.6 + 49.4 + 0.6) * 1e6,
                      traffic_source='profiles/cms_scale_kernels_r01.txt: k_along_msc_limit + '
                                     'k_along_propagate_field + k_along_msc_apply + k_along_finish '
                                     'at one saturated iteration (ncu --set full)',
                      events=100, per_event=10,
                      label='CMS-scale stand-in geometry (tools/make_cms_scale.py: 4 levels, 2916 '
                            'unit volumes, 2 rect arrays, BIH), 1 T uniform field, full EM, '
                            '%d x %d isotropic 10 GeV e-/gamma primaries from the origin per GPU'),
    # BASELINE configs[2]: simple-CMS nested cylinders, 1 T uniform field, 10 GeV e-/gamma mix
    'simple-cms': dict(image=os.path.join(REPO, 'data', 'images',
                                          'simple-cms-em-field-initcharge.b2img'),
                       config=os.path.join(REPO, 'data', 'images', 'simple-cms-em-field.json'),
                       alg_bytes=ALG_BYTES_PER_TRACK_STEP, events=100, per_event=10,
                       traffic=None, traffic_source='no ncu capture for this workload',
                       label='simple-CMS nested cylinders (test/geocel/data/simple-cms.org.json), '
                             '1 T uniform field, full EM, %d x %d isotropic 10 GeV e-/gamma '
                             'primaries from the origin per GPU'),
}


# The CMS-scale problem at 1000 x 10 primaries per GPU: the track slots saturate and the pass is
# throughput-bound; at 100 x 10 half of the pass is the latency of looping-track chains, whose
# length differs from one rank's events to another's (profiles/README_r02.md)
WORKLOADS['cms-scale-10k'] = dict(WORKLOADS['cms-scale'], events=1000, per_event=10)


def make_workload_events(workload, params_or_problem, num_events, per_event, first_event, dtype):
    """Primaries of one rank: events [first_event, first_event + num_events)."""
    if workload == 'testem3':
        return make_events(num_events, per_event, first_event, 1, dtype)
    opts = {'seed': 20220904, 'pdg': [11, 22], 'num_events': first_event + num_events,
            'primaries_per_event': per_event, 'energy': 10000.0, 'position': [0, 0, 0],
            'direction': {'distribution': 'isotropic'}}
    out = params_or_problem.generate_primaries(opts)
    prim = out[0] if isinstance(out, tuple) else out
    prim = np.ascontiguousarray(prim[first_event * per_event:]).astype(dtype)
    offsets = np.arange(0, len(prim) + 1, per_event, dtype=np.uint32)
    return prim, offsets


def make_events(num_events, per_event, first_event, particle_id, dtype):
    n = num_events * per_event
    p = np.zeros(n, dtype=dtype)
    p['particle_id'] = particle_id
    p['energy'] = ENERGY_MEV
    p['pos'] = (-22, 0, 0)
    p['dir'] = (1, 0, 0)
    p['event_id'] = first_event + np.repeat(np.arange(num_events), per_event)
    offsets = np.arange(0, n + 1, per_event, dtype=np.uint32)
    return p, offsets


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region (B200_PROFILING.md
    clocks line). Sampled in-process through NVML, the library behind nvidia-smi: forking
    nvidia-smi every 200 ms stalled the launching threads for tens of milliseconds and
    showed up in the end-to-end number. nvidia-smi is the fallback if NVML cannot load."""
    REASONS = {0x8: 'hw_slowdown', 0x40: 'hw_thermal_slowdown', 0x20: 'sw_thermal_slowdown',
               0x4: 'sw_power_cap'}

    def __init__(self, index, interval=None):
        # NVML queries go through the driver: on an 8-GPU box every rank polling at 20 Hz is
        # 300 driver calls per second next to a million kernel launches. Rank 0 (whose record
        # is printed) keeps 50 ms; the other ranks poll at 200 ms
        if interval is None:
            interval = 0.05 if int(os.environ.get('RANK', '0')) == 0 else 0.2
        self.interval = interval
        self.samples = []
        self.reasons = set()
        self.max_mhz = None
        self._stop = threading.Event()
        self.index = index
        self._thread = threading.Thread(target=self._run, daemon=True)
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            # NVML enumerates physical devices: map through CUDA_VISIBLE_DEVICES
            visible = os.environ.get('CUDA_VISIBLE_DEVICES')
            phys = index
            if visible:
                ids = [v.strip() for v in visible.split(',') if v.strip()]
                if index < len(ids) and ids[index].isdigit():
                    phys = int(ids[index])
            self._handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self._handle,
                                                                  pynvml.NVML_CLOCK_SM))
            self._nvml = pynvml
        except Exception:
            self._nvml = None

    def _sample_nvml(self):
        n = self._nvml
        self.samples.append(float(n.nvmlDeviceGetClockInfo(self._handle, n.NVML_CLOCK_SM)))
        try:
            mask = n.nvmlDeviceGetCurrentClocksEventReasons(self._handle)
        except Exception:
            mask = n.nvmlDeviceGetCurrentClocksThrottleReasons(self._handle)
        for bit, name in self.REASONS.items():
            if mask & bit:
                self.reasons.add(name)

    def _sample_smi(self):
        q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,'
             'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
             'clocks_event_reasons.sw_power_cap')
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + q,
                              '--format=csv,noheader,nounits'],
                             capture_output=True, text=True, timeout=5).stdout.strip()
        f = [x.strip() for x in out.split(',')]
        self.samples.append(float(f[0]))
        self.max_mhz = float(f[1])
        for nm, v in zip(names, f[2:]):
            if v.lower().startswith('active'):
                self.reasons.add(nm)

    def _run(self):
        while not self._stop.is_set():
            try:
                if self._nvml is not None:
                    self._sample_nvml()
                else:
                    self._sample_smi()
            except Exception:
                pass
            self._stop.wait(self.interval if self._nvml is not None else 0.5)

    def __enter__(self):
        self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thread.join(timeout=5)

    def summary(self):
        return {'sm_mhz': float(np.median(self.samples)) if self.samples else None,
                'sm_max_mhz': self.max_mhz, 'reasons': sorted(self.reasons),
                'source': 'nvml' if self._nvml is not None else 'nvidia-smi',
                'samples': len(self.samples)}


def measured_peak_gbs():
    path = os.path.join(REPO, 'MEASURED_PEAKS.json')
    if os.path.exists(path):
        return json.load(open(path))['hbm_gbs'], 'measured'
    return 6650.0, 'fallback'


def cpu_reference_run(num_events, per_event, slots_per_stream, threads, workload='testem3'):
    """Time the reference's own host Stepper (one Stepper per OpenMP thread)."""
    sys.path.insert(0, os.path.join(REPO, 'oracle'))
    import celerref
    cfg = json.load(open(WORKLOADS[workload]['config']))
    cfg['max_streams'] = max(threads, 1)
    cfg['initializer_capacity'] = 1 << 22
    cfg['track_order'] = 'none'  # the reference's CPU default
    problem = celerref.Problem(cfg)
    # particle id of e- is fixed by the physics file order (e+, e-, gamma)
    prim, offsets = make_workload_events(workload, problem, num_events, per_event, 0,
                                         celerref.PRIMARY_DTYPE)
    r = problem.run_events(prim, offsets, slots_per_stream, threads)
    return r


def run_reference_arm(args, rank, world):
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    threads = cores
    per_event, num_events = 4, max(2 * threads, 8)
    if args.workload != 'testem3':
        per_event, num_events = 1, max(threads, 4)  # 10 GeV showers: ~10x the steps each
    # warm-up passes are smaller; each timed step is the same bounded sample
    for _ in range(args.warmup):
        cpu_reference_run(max(threads, 4), 1, 4096, threads, args.workload)
    steps, secs = 0, 0.0
    iters = 0
    for _ in range(args.steps):
        r = cpu_reference_run(num_events, per_event, 4096, threads, args.workload)
        steps += r['num_steps']
        secs += r['seconds']
        iters += r['num_step_iterations']
    value = steps / secs
    sample = ('%d events x %d primaries per step on %d OpenMP threads, one Stepper '
              'per thread, 4096 track slots each' % (num_events, per_event, threads))
    line = {'impl': 'reference', 'metric': 'track-steps/sec', 'value': value,
            'unit': 'track-steps/s', 'n_gpus': args.gpus, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': 1e3 * secs / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64',
            'data': 'synthetic',
            'config': {'workload': (WORKLOADS[args.workload]['label'] % (num_events, per_event))
                                   + '; steel/lAr stand-in physics', 'sample': sample},
            'cpu_baseline': {'value': value, 'unit': 'track-steps/s', 'cores': threads,
                             'kind': 'reference', 'sample': sample},
            'e2e': {'value': value, 'unit': 'track-steps/s', 'h2d_bytes_per_step': 0,
                    'd2h_bytes_per_step': 0},
            'events_per_sec': num_events * args.steps / secs}
    print(json.dumps(line))


def run_reference_cuda_arm(args, rank, world):
    """The reference's OWN CUDA build (oracle/_ref/libcelerref_cuda.so: its .cu files compiled
    for sm_100 by oracle/Makefile with the reference's CMake defaults, CELERITAS_MAX_BLOCK_SIZE
    256) on this GPU, same workload, same primaries, same number of track slots, events merged
    onto one state (celer-sim's GPU configuration: one stream, track_order init_charge). Not the
    driver's reference arm (that is the CPU one): this is the number BASELINE.json's target is
    stated against. SM clocks are sampled during the timed region like in the main arm."""
    if rank != 0:
        return
    os.environ['CELERREF_CUDA'] = '1'
    sys.path.insert(0, os.path.join(REPO, 'oracle'))
    import celerref
    wl = WORKLOADS[args.workload]
    events = args.events or wl['events']
    per_event = args.primaries_per_event or wl['per_event']
    cfg = json.load(open(wl['config']))
    cfg['track_order'] = 'init_charge'
    problem = celerref.Problem(cfg)
    prim, _ = make_workload_events(args.workload, problem, events, per_event, 0,
                                   celerref.PRIMARY_DTYPE)
    for _ in range(args.warmup):
        celerref.run_merged_device(problem, prim, args.slots)
    steps, secs, iters = 0, 0.0, 0
    with ClockSampler(int(os.environ.get('LOCAL_RANK', '0'))) as clocks:
        for _ in range(args.steps):
            r = celerref.run_merged_device(problem, prim, args.slots)
            steps += r['num_steps']
            secs += r['seconds']
            iters += r['num_step_iterations']
    line = {'impl': 'reference-cuda', 'metric': 'track-steps/sec', 'value': steps / secs,
            'unit': 'track-steps/s', 'n_gpus': 1, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': 1e3 * secs / args.steps, 'higher_is_better': True, 'dtype': 'f64',
            'data': 'synthetic', 'num_step_iterations': iters,
            'track_steps_per_pass': steps // max(args.steps, 1),
            'events_per_sec': events * args.steps / secs,
            'clocks': clocks.summary(),
            'config': {'workload': (wl['label'] % (events, per_event))
                                   + ', merged events, %d track slots on one stream, track_order '
                                     'init_charge; steel/lAr stand-in physics' % args.slots,
                       'build': "the reference's own .cu sources, nvcc -O3 -gencode "
                                'arch=compute_100,code=sm_100, CELERITAS_MAX_BLOCK_SIZE 256 '
                                '(oracle/Makefile ref_cuda)'}}
    print(json.dumps(line))


def reference_cuda_subrecord(workload, slots, events, per_event, b200_value, steps=2, warmup=1):
    """Run the reference's own CUDA build on the same GPU right after our timed region, in a
    child process (its library and ours both own a CUDA context; the child keeps the CPU
    oracle library out of this process), and return {value, ms_per_step, clocks, ratio}."""
    lib = os.path.join(REPO, 'oracle', '_ref', 'libcelerref_cuda.so')
    if not os.path.exists(lib):
        return {'unavailable': 'oracle/_ref/libcelerref_cuda.so not built (make -C oracle ref_cuda)'}
    cmd = [sys.executable, os.path.abspath(__file__), '--impl', 'reference-cuda',
           '--workload', workload, '--steps', str(steps), '--warmup', str(warmup),
           '--slots', str(slots), '--events', str(events),
           '--primaries-per-event', str(per_event)]
    env = dict(os.environ)
    for k in ('RANK', 'LOCAL_RANK', 'WORLD_SIZE'):
        env.pop(k, None)
    try:
        out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    except subprocess.TimeoutExpired:
        return {'unavailable': 'reference CUDA build timed out after 600 s'}
    for text in reversed(out.stdout.strip().splitlines()):
        if text.startswith('{'):
            r = json.loads(text)
            return {'value': r['value'], 'unit': r['unit'], 'ms_per_step': r['ms_per_step'],
                    'steps': r['steps'], 'warmup': r['warmup'],
                    'num_step_iterations': r['num_step_iterations'],
                    'track_steps_per_pass': r['track_steps_per_pass'],
                    'clocks': r['clocks'], 'ratio': b200_value / r['value'],
                    'ratio_is': 'this library (value) / reference CUDA build (value), same GPU, '
                                'same workload and primaries, measured back to back',
                    'build': r['config']['build']}
    return {'unavailable': 'reference CUDA arm failed: ' + (out.stderr.strip()[-300:] or 'no output')}


# Algorithmic bytes of ONE action per track-step ("byte shares"): the
# per-slot fields that action has to read plus the ones it has to write, at D levels and P
# processes, out of the 2 * S_live of SURVEY.md 8(d).
def action_alg_bytes(action, depth, nproc):
    geo = 31 + 56 * depth
    if action.startswith('along-step'):
        # reads geo, particle 12, material 4, sim 45, phys scalars 48 + msc 32, rng 24;
        # writes geo, sim 45, energy 8, phys 40 + msc 32, rng 24
        return (geo + 12 + 4 + 45 + 48 + 32 + 24) + (geo + 45 + 8 + 40 + 32 + 24)
    if action == 'pre-step':
        return (12 + 4 + 45 + 8 + 24 + 8 + 4 * depth) + (45 + 8 + 8 * nproc + 16 + 24 + 16)
    return None


def run_b200(args, workload, steps, warmup, rank, local_rank, world, dist, events, per_event,
             with_roofline=True):
    """One workload on this rank's GPU: returns the measurements of the timed regions."""
    import torch
    import celeritas_b200 as cb
    wl = WORKLOADS[workload]
    params = cb.Params(os.environ.get('B200_BENCH_IMAGE', wl['image'])
                       if workload == 'testem3' else wl['image'])
    nstreams = max(args.streams, 1)
    steppers = [cb.Stepper(params, args.slots // nstreams, stream_id=rank * nstreams + k)
                for k in range(nstreams)]
    assert params.find_particle(11) == 1
    ndet = params.num_detectors
    # Events are sharded by rank: rank r owns global events [r*E, (r+1)*E)
    prim, offsets = make_workload_events(workload, params, events, per_event, rank * events,
                                         cb.PRIMARY_DTYPE)
    nprim = len(prim)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def one_pass():
        for st in steppers:
            st.calo_clear()
        if nstreams == 1:
            return steppers[0].run_events(prim, offsets, merge_events=True)
        per_stream, seconds = cb.run_events_streams(steppers, prim, offsets, merge_events=True)
        r = {k: sum(x[k] for x in per_stream)
             for k in ('num_steps', 'num_primaries', 'num_tracks', 'num_aborted')}
        # the longest stream sets the number of step iterations of the pass
        r['num_step_iterations'] = max(x['num_step_iterations'] for x in per_stream)
        r['seconds'] = seconds
        return r

    def reduce_tallies(r):
        """End-of-run reduction of tallies and counters over NVLink (NCCL)."""
        calo = torch.from_numpy(sum(st.calo() for st in steppers)).cuda()
        counts = torch.tensor([r['num_steps'], r['num_step_iterations'], r['num_primaries']],
                              dtype=torch.int64, device='cuda')
        if dist is not None:
            dist.all_reduce(calo)
            dist.all_reduce(counts)
        return calo.cpu().numpy(), counts.cpu().numpy()

    for _ in range(warmup):
        reduce_tallies(one_pass())

    # ---- timed region 1: e2e through the public C-ABI with HOST buffers
    launches0 = cb.launch_count()
    barrier()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(local_rank) as clocks:
        ev0.record()
        total_steps = total_iters = 0
        dev_secs = 0.0
        for _ in range(steps):
            r = one_pass()
            calo, counts = reduce_tallies(r)
            total_steps += int(counts[0])
            total_iters += int(counts[1])
            dev_secs += r['seconds']
        ev1.record()
        barrier()
    launches = cb.launch_count() - launches0
    e2e_secs = torch.tensor([ev0.elapsed_time(ev1) * 1e-3], device='cuda', dtype=torch.float64)
    dsecs = torch.tensor([dev_secs], device='cuda', dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(e2e_secs, op=dist.ReduceOp.MAX)
        dist.all_reduce(dsecs, op=dist.ReduceOp.MAX)
    e2e_secs = float(e2e_secs.item())
    dsecs = float(dsecs.item())
    value = total_steps / dsecs
    depth = params.max_depth
    out = {
        'value': value, 'unit': 'track-steps/s', 'ms_per_step': 1e3 * dsecs / steps,
        'steps': steps, 'warmup': warmup,
        'events_per_sec': events * steps * world / dsecs,
        'num_step_iterations': total_iters,
        'track_steps_per_pass_per_gpu': total_steps // max(steps * world, 1),
        'e2e': {'value': total_steps / e2e_secs, 'unit': 'track-steps/s',
                'h2d_bytes_per_step': int(nprim * 72 + 12 * nprim),
                'd2h_bytes_per_step': int(64 * (total_iters // max(steps * world, 1))
                                          + 8 * ndet)},
        'gpu_launches': int(launches),
        'clocks': clocks.summary(),
        'workload': (wl['label'] % (events, per_event))
                    + (', merged events, %d track slots over %d concurrent stream(s), '
                       'track_order init_charge; steel/lAr stand-in physics '
                       '(tools/make_physics.py)' % (args.slots, nstreams)),
        'calo_sum_mev': float(calo.sum()),
    }
    if with_roofline:
        # ---- timed region 2: per-action CUDA-event timing (one stream holding the whole
        # workload, one launch per action) for the roofline of the dominant kernel
        stepper = steppers[0] if nstreams == 1 else cb.Stepper(params, args.slots,
                                                               stream_id=rank * nstreams)
        stepper.set_action_times(True)
        before = stepper.action_times
        stepper.calo_clear()
        r2 = stepper.run_events(prim, offsets, merge_events=True)
        after = stepper.action_times
        stepper.set_action_times(False)
        per_action = {k: after[k] - before.get(k, 0.0) for k in after}
        top = max(per_action, key=per_action.get)
        top_secs = per_action[top]
        total_action_secs = sum(per_action.values())
        peak, peak_kind = measured_peak_gbs()
        nproc = 4
        alg_step = 2 * (248 + 56 * depth + 8 * nproc)
        alg_action = action_alg_bytes(top, depth, nproc) or alg_step
        # every track-step of the pass goes through this action exactly once
        achieved = r2['num_steps'] * alg_action / top_secs / 1e9
        out['roofline'] = {
            'bound': 'hbm', 'kernel': top, 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
            'frac': achieved / peak, 'traffic': wl['traffic'],
            'traffic_source': wl['traffic_source'], 'peak_kind': peak_kind,
            'alg_bytes_per_track_step_this_action': alg_action,
            'alg_bytes_per_track_step_whole_step': alg_step,
            'kernel_share_of_step': top_secs / total_action_secs,
            # track-steps/s x B_alg over the peak of ALL GPUs of the job
            'whole_step': {'achieved': value * alg_step / 1e9,
                           'peak': peak * world,
                           'frac': value * alg_step / 1e9 / (peak * world)},
            'per_action_seconds': per_action}
    del steppers
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=3)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference', 'reference-cuda'])
    ap.add_argument('--workload', default='testem3', choices=sorted(WORKLOADS))
    ap.add_argument('--events', type=int, default=None)
    ap.add_argument('--primaries-per-event', type=int, default=None)
    ap.add_argument('--slots', type=int, default=NUM_TRACK_SLOTS)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-extra', action='store_true',
                    help='skip the reference-CUDA sub-record, the extra workloads and the '
                         'strong-scaling pass (kernel tuning runs)')
    ap.add_argument('--streams', type=int, default=NUM_STREAMS,
                    help='concurrent steppers (CUDA streams) per GPU; slots are divided among them')
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    wl = WORKLOADS[args.workload]
    if args.events is None:
        args.events = wl['events']
    if args.primaries_per_event is None:
        args.primaries_per_event = wl['per_event']

    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    world = int(os.environ.get('WORLD_SIZE', '1'))

    if args.impl == 'reference':
        run_reference_arm(args, rank, world)
        return
    if args.impl == 'reference-cuda':
        run_reference_cuda_arm(args, rank, world)
        return

    import torch
    import celeritas_b200 as cb
    torch.cuda.set_device(local_rank)
    cb.set_device(local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    main_run = run_b200(args, args.workload, args.steps, args.warmup, rank, local_rank, world,
                        dist, args.events, args.primaries_per_event)

    # The other workloads of BASELINE.json (configs[2] simple-CMS, configs[3]/[4] CMS-scale)
    # on every rank (weak scaling: every rank its own events), three timed passes each
    extra = {}
    strong = None
    if not args.no_extra and args.workload == 'testem3':
        for name in ('cms-scale', 'cms-scale-10k', 'simple-cms'):
            if name == 'simple-cms' and world > 1:
                continue  # configs[4] is the CMS-scale sweep
            w = WORKLOADS[name]
            big = name == 'cms-scale-10k'
            extra[name] = run_b200(args, name, 2 if big else 3, 3, rank, local_rank, world, dist,
                                   w['events'], w['per_event'],
                                   with_roofline=(world == 1 and not big))
        if world > 1 and args.events % world == 0:
            # strong scaling: the SAME 100 x 100 primaries split over the N GPUs
            strong = run_b200(args, args.workload, 3, 3, rank, local_rank, world, dist,
                              args.events // world, args.primaries_per_event,
                              with_roofline=False)

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    r = main_run
    line = {
        'metric': 'track-steps/sec', 'value': r['value'], 'unit': 'track-steps/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': r['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
        'config': {'workload': r['workload'],
                   'l2': 'working set %.0f MB of SoA state per pass exceeds the 126 MB L2'
                         % (args.slots * 336 / 1e6),
                   'parallelism': 'events sharded by rank, NCCL all-reduce of tallies'},
        'events_per_sec': r['events_per_sec'],
        'num_step_iterations': r['num_step_iterations'],
        'e2e': r['e2e'], 'gpu_launches': r['gpu_launches'], 'clocks': r['clocks'],
        'roofline': r['roofline'],
    }
    if strong is not None:
        line['strong_scaling'] = {
            'what': 'the same %d x %d primaries split over %d GPUs (events sharded by rank)'
                    % (args.events, args.primaries_per_event, world),
            'value': strong['value'], 'unit': 'track-steps/s',
            'events_per_sec': strong['events_per_sec'], 'ms_per_step': strong['ms_per_step'],
            'e2e': strong['e2e'], 'clocks': strong['clocks']}
    if not args.no_extra and world == 1:
        line['reference_cuda'] = reference_cuda_subrecord(
            args.workload, args.slots, args.events, args.primaries_per_event, r['value'])
    if extra:
        line['workloads'] = {}
        for name, x in extra.items():
            rec = {'b200': {k: x[k] for k in ('value', 'unit', 'ms_per_step', 'steps', 'warmup',
                                              'events_per_sec', 'num_step_iterations', 'e2e',
                                              'gpu_launches', 'track_steps_per_pass_per_gpu')},
                   'clocks': x['clocks'], 'config': {'workload': x['workload']},
                   'n_gpus': world, 'scaling': 'weak'}
            if 'roofline' in x:
                rec['roofline'] = x['roofline']
            if world == 1:
                w = WORKLOADS[name]
                rec['reference_cuda'] = reference_cuda_subrecord(
                    name, args.slots, w['events'], w['per_event'], x['value'])
            line['workloads'][name] = rec
    if not args.no_cpu_baseline and world == 1:
        cores = os.cpu_count() or 1
        ne, pe = max(2 * cores, 8), 4
        if args.workload != 'testem3':
            ne, pe = max(cores, 4), 1
        rc = cpu_reference_run(ne, pe, 4096, cores, args.workload)
        line['cpu_baseline'] = {
            'value': rc['num_steps'] / rc['seconds'], 'unit': 'track-steps/s', 'cores': cores,
            'kind': 'reference',
            'sample': '%d events x %d primaries of the same workload (a bounded sample: the GPU '
                      'arm runs %d x %d merged onto one state), reference host Stepper '
                      '(oracle/_ref), one per OpenMP thread, 4096 slots each, track_order none'
                      % (ne, pe, args.events, args.primaries_per_event)}
    print(json.dumps(line))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
