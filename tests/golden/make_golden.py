#!/usr/bin/env python
"""Generate tests/golden/control_plane.json by running the UNMODIFIED reference
control plane (/root/reference/tfmesos) under the three stubs in
oracle/refstubs (pymesos, addict, empty tensorflow) -- SURVEY.md 4.3 / 8c.

Run in the build container only (the GPU box has no /root/reference):

    python tests/golden/make_golden.py

Everything random or host-specific is masked before it is written: task uuids
-> "<id>", the scheduler address -> "<sched>", the interpreter -> "<python>",
task addresses -> "<job:idx>", PYTHONPATH's value -> "<pythonpath>", cwd ->
"<cwd>".  The product's tests (tests/test_control_plane_parity.py) drive
tfmesos_b200 through the same scripts and compare after the same masking.
"""
import json
import os
import socket
import subprocess
import sys
import tempfile
import threading

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from tests.golden import scenarios  # noqa: E402

REF = "/root/reference"
STUBS = os.path.join(ROOT, "oracle", "refstubs")


def load_reference():
    sys.path.insert(0, STUBS)
    sys.path.insert(0, REF)
    import tfmesos  # noqa: F401  (the reference package)
    import tfmesos.scheduler as rs
    assert rs.__file__.startswith(REF), rs.__file__
    return tfmesos, rs


def main():
    tfmesos, rs = load_reference()
    import addict
    os.environ.pop("DOCKER_IMAGE", None)

    out = {"reference": "douban/tfmesos @ /root/reference (py3 insertion-order "
                        "task iteration, SURVEY.md 8c)"}
    out["placement"] = scenarios.run_placement(
        make_scheduler=lambda jobs, **kw: rs.TFMesosScheduler(
            [rs.Job(**j) for j in jobs], master="stub", quiet=True, **kw),
        make_offer=lambda spec: scenarios.build_offer(addict.Dict, spec))
    out["cluster_def"] = scenarios.run_cluster_def(
        make_scheduler=lambda jobs, **kw: rs.TFMesosScheduler(
            [rs.Job(**j) for j in jobs], master="stub", quiet=True, **kw))
    out["status"] = scenarios.run_status(
        make_scheduler=lambda jobs, **kw: rs.TFMesosScheduler(
            [rs.Job(**j) for j in jobs], master="stub", quiet=True, **kw),
        make_update=lambda **kw: addict.Dict(**kw))
    out["job_normalisation"] = scenarios.run_job_normalisation(
        cluster=tfmesos.cluster, Job=rs.Job, sched_module=rs)
    out["replica_mode"] = run_replica_mode(rs, addict)
    out["wire"] = scenarios.run_wire(__import__("tfmesos.utils", fromlist=["x"]))

    path = os.path.join(HERE, "control_plane.json")
    with open(path, "w") as f:
        json.dump(out, f, indent=1, sort_keys=True)
        f.write("\n")
    print("wrote", path)


def run_replica_mode(rs, addict):
    """Whole replica-mode path of the reference: scheduler.start() with a fake
    driver that Popen()s `python -m tfmesos.server <id> <addr>` for every
    TaskInfo (scheduler.py:162-176 -> server.py:14-113)."""
    results = {}
    tmp = tempfile.mkdtemp(prefix="tfmesos_golden_")
    probe = os.path.join(HERE, "probe_child.py")
    cmd = ("%s %s %s --ps_hosts {ps_hosts} --worker_hosts {worker_hosts} "
           "--job_name {job_name} --worker_index {task_index}"
           % (sys.executable, probe, tmp))
    jobs = [dict(name="ps", num=1, cmd=cmd),
            dict(name="worker", num=2, cmd=cmd, gpus=1)]
    procs = []
    captured = {}

    class FakeDriver(object):
        version = "1.0.0"

        def __init__(self, sched, framework, master, use_addict=False):
            self.sched = sched
            captured["framework_keys"] = sorted(framework.keys())

        def start(self):
            fid = addict.Dict(value="fw-1")
            mi = addict.Dict(hostname="localhost", port=5050)
            self.sched.registered(self, fid, mi)
            offer = scenarios.build_offer(addict.Dict, dict(
                id="o1", cpus=8.0, mem=16384.0, gpus=["GPU-0", "GPU-1", "GPU-2"],
                gpu_type="SET"))
            self.sched.resourceOffers(self, [offer])

        def launchTasks(self, offer_id, infos):
            for ti in infos:
                env = dict(os.environ)
                for var in ti.command.environment.variables:
                    env[var.name] = var.value
                env["PYTHONPATH"] = os.pathsep.join(
                    [STUBS, REF, env.get("PYTHONPATH", "")])
                # neutral cwd: the repo root holds a `tfmesos` alias package that
                # would shadow the reference for `python -m tfmesos.server`
                p = subprocess.Popen(ti.command.value, shell=True, env=env, cwd=tmp,
                                     stdout=subprocess.DEVNULL,
                                     stderr=subprocess.DEVNULL)
                procs.append((ti.task_id.value, p))

        def suppressOffers(self):
            pass

        def declineOffer(self, *a):
            pass

        def reviveOffers(self):
            pass

        def stop(self):
            pass

        def join(self):
            pass

    rs.MesosSchedulerDriver = FakeDriver
    hooks = os.path.join(tmp, "hooks.txt")
    extra = {"initializer": "echo init >> %s" % hooks, "finalizer": "echo fin >> %s" % hooks}
    sched = rs.TFMesosScheduler([rs.Job(**j) for j in jobs], master="stub",
                                quiet=True, extra_config=extra,
                                forward_addresses=None)
    # the reference's start() waits for ever if a task never registers
    watchdog = threading.Timer(120, lambda: os._exit(3))
    watchdog.daemon = True
    watchdog.start()
    sched.start()
    watchdog.cancel()
    names = {t.mesos_task_id: (t.job_name, t.task_index, t.addr)
             for t in sched.tasks.values()}
    addr_mask = {a: "<%s:%s>" % (j, i) for (j, i, a) in names.values()}
    results["targets"] = {k: scenarios.mask_addr(v, addr_mask)
                          for k, v in sched.targets.items()}
    results["started"] = sched.started
    results["framework_keys"] = captured["framework_keys"]
    for tid, p in procs:
        rc = p.wait(timeout=60)
        upd = addict.Dict()
        upd.task_id.value = tid
        upd.state = "TASK_FINISHED" if rc == 0 else "TASK_FAILED"
        upd.message = ""
        sched.statusUpdate(None, upd)
    results["finished"] = sched.finished()
    results["job_finished"] = dict(sched.job_finished)
    with open(hooks) as f:
        lines = f.read().split()
    # extra_config initializer / finalizer (server.py:68-70,106-109)
    results["extra_config_hooks"] = {"init": lines.count("init"), "fin": lines.count("fin")}
    children = {}
    for fn in sorted(os.listdir(tmp)):
        if not fn.endswith(".json"):
            continue
        with open(os.path.join(tmp, fn)) as f:
            rec = json.load(f)
        rec["argv"] = [scenarios.mask_addr(a, addr_mask) for a in rec["argv"]]
        rec["env"] = {k: scenarios.mask_addr(v, addr_mask)
                      for k, v in rec["env"].items()}
        rec["cwd"] = "<cwd>" if rec["cwd"] == os.getcwd() else rec["cwd"]
        children[fn[:-5]] = rec
    results["children"] = children
    sched.stop()
    return results


if __name__ == "__main__":
    main()
