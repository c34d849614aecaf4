"""torchrun check of the sharded populate (run with 2+ GPUs):
  * the shared-host-pool path and the all-gather path return the same set of records
    (same seed; only the order differs: turn-major vs rank-major), identical on every rank;
  * end-to-end time per populate of both paths."""
import hashlib, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
import bench
from nessai_b200.livepoint import numpy_array_to_live_points
from nessai_b200.proposal import B200FlowProposal

rank, world, lr = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(lr)
dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
g, cfg, sd = bench.load_c2()
live, _ = bench.live_points()
model = bench.GaussianModel()
pool = int(os.environ.get("POOL", 1_000_000)) * world


def make():
    torch.manual_seed(1)
    p = B200FlowProposal(model, rng=np.random.default_rng(1), flow_config=cfg,
                         training_config=dict(device_tag=f"cuda:{lr}"), output=tempfile.mkdtemp(),
                         poolsize=pool, drawsize=pool, device_prior="auto")
    p.initialise()
    ls = numpy_array_to_live_points(live, model.names)
    ls["logL"] = model.log_likelihood(ls)
    p.check_state(ls)
    p.flow.model.load_state_dict({k: torch.from_numpy(v) for k, v in sd.items()})
    p.flow.model.eval()
    return p, ls[0]


def digest(a):
    return hashlib.sha1(np.ascontiguousarray(a).view(np.uint8).tobytes()).hexdigest()


res = {}
for name in ("shared", "allgather"):
    p, worst = make()
    eng = p._get_engine()
    eng.seed = 4242
    if name == "allgather":
        eng._pool_ok, eng._pool, eng._pool_turn = False, None, 0
    for _ in range(3):
        p.populate(worst, n_samples=pool, max_samples=pool)
    eng._turn_rows = 10 * pool  # same Philox window for both paths
    dist.barrier(); torch.cuda.synchronize()
    p.population_time *= 0
    reps = 8
    for _ in range(reps):
        p.populate(worst, n_samples=pool, max_samples=pool)
    t = p.population_time.total_seconds() / reps
    tr = getattr(eng, "last_trace", None)
    if tr and os.environ.get("NB200_TRACE"):
        line = " ".join(f"{b}+{1e3 * (tb - ta):.3f}" for (a, ta), (b, tb) in zip(tr[:-1], tr[1:]))
        print(f"[trace {name} rank {rank}] {line}", flush=True)
    eng._turn_rows = 10 * pool
    p.populate(worst, n_samples=pool, max_samples=pool)
    rows = np.stack([p.samples[n] for n in model.names] + [p.samples["logP"]], axis=1)  # plain (n, D + 1) float64
    hs = [None] * world
    dist.all_gather_object(hs, digest(rows))
    assert len(set(hs)) == 1, f"{name}: ranks disagree on the pool"
    order = np.lexsort([rows[:, 2], rows[:, 1], rows[:, 0]])
    res[name] = (digest(rows[order]), len(rows), t, p.n_proposed)
    if rank == 0:
        print(f"{name:9s}: {len(rows)} records, {1e3 * t:.3f} ms per populate of {p.n_proposed} rows "
              f"= {p.n_proposed / t:.3e} rows/s; pool registered: {getattr(eng._pool, 'registered', None)}", flush=True)
if rank == 0:
    same = res["shared"][:2] == res["allgather"][:2]
    print("same record set through both paths:", same, flush=True)
    assert same
dist.destroy_process_group()
