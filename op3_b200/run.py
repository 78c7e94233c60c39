"""Run an UNMODIFIED reference script on top of the sm_100a kernels:

    python -m op3_b200.run [--precision f16x3|fp32|tf32x3|tf32] <script.py> [script args...]
    torchrun --nproc-per-node 8 -m op3_b200.run exps/train_op3/train_op3.py -va pickplace -de 0

`install()` aliases the `op3.torch.*` module names to this package, then the script runs as `__main__` through runpy --
no edit of exps/train_op3/train_op3.py (:78-116) or the MPC scripts is needed.  Under torchrun each process is pinned to
its own GPU *before* CUDA is initialised (CUDA_VISIBLE_DEVICES = LOCAL_RANK), so the script's
`nn.DataParallel(m)` (train_op3.py:99-100) wraps a single device and degenerates to a plain call, and the NCCL process
group is created here: OP3Trainer.train_step then all-reduces the flat gradient (one process per GPU replaces the
reference's intra-process replication).
"""
import os
import runpy
import sys


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    precision = os.environ.get("OP3_PRECISION", "f16x3")
    while argv and argv[0].startswith("--"):
        if argv[0] == "--precision" and len(argv) > 1:
            precision = argv[1]
            argv = argv[2:]
        elif argv[0].startswith("--precision="):
            precision = argv[0].split("=", 1)[1]
            argv = argv[1:]
        else:
            raise SystemExit("op3_b200.run: unknown option %s" % argv[0])
    if not argv:
        raise SystemExit(__doc__)
    script = os.path.abspath(argv[0])
    if not os.path.exists(script):
        raise SystemExit("op3_b200.run: %s does not exist" % script)

    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world > 1 and "LOCAL_RANK" in os.environ:
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        ids = visible.split(",") if visible else None
        lr = int(os.environ["LOCAL_RANK"])
        os.environ["CUDA_VISIBLE_DEVICES"] = ids[lr] if ids else str(lr)      # before the first CUDA call

    # the reference checkout the script belongs to must be importable (`import op3...`, `import exps...`)
    d = os.path.dirname(script)
    while d and d != os.path.dirname(d):
        if os.path.isdir(os.path.join(d, "op3")) and os.path.isdir(os.path.join(d, "exps")):
            if d not in sys.path:
                sys.path.insert(0, d)
            break
        d = os.path.dirname(d)
    import op3_b200
    op3_b200.install()
    op3_b200.set_precision(precision)

    if world > 1:
        import torch
        import torch.distributed as dist
        if not dist.is_initialized():
            backend = "nccl" if torch.cuda.is_available() else "gloo"
            dist.init_process_group(backend)

    sys.argv = [script] + argv[1:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
