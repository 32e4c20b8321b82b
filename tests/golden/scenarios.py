"""Scripts that drive a tfmesos-style scheduler through its PUBLIC callbacks and
record what it did, with everything random masked.  The same scripts are run
against the unmodified reference (tests/golden/make_golden.py -> the committed
control_plane.json) and against tfmesos_b200 (tests/test_control_plane_parity.py).

Reference behaviour exercised: tfmesos/scheduler.py:183-221 (task enumeration),
:223-277 (first-fit offers), :279-318 (targets, cluster_def, config dict),
:384-457 (failure policy), :474-477 (finished); tfmesos/__init__.py:7-22;
tfmesos/utils.py:6-15.
"""
import pickle
import socket
import struct
import sys
import threading

MASKED_ENV_VALUES = {"PYTHONPATH": "<pythonpath>"}


# ----------------------------------------------------------------------------
def build_offer(D, spec):
    """spec: {id, cpus, mem, gpus (list for SET | int for SCALAR | None), gpu_type}"""
    offer = D()
    offer.id.value = spec["id"]
    offer.agent_id.value = spec.get("agent", "agent-1")
    offer.hostname = spec.get("hostname", "localhost")
    res = []
    if spec.get("cpus") is not None:
        r = D()
        r.name, r.type = "cpus", "SCALAR"
        r.scalar.value = spec["cpus"]
        res.append(r)
    if spec.get("mem") is not None:
        r = D()
        r.name, r.type = "mem", "SCALAR"
        r.scalar.value = spec["mem"]
        res.append(r)
    if spec.get("gpus") is not None:
        r = D()
        r.name = "gpus"
        r.type = spec.get("gpu_type", "SET")
        if r.type == "SET":
            r.set.item = list(spec["gpus"])
        else:
            r.scalar.value = spec["gpus"]
        res.append(r)
    offer.resources = res
    return offer


def mask_addr(s, addr_mask):
    if not isinstance(s, str):
        return s
    for a, m in sorted(addr_mask.items(), key=lambda kv: -len(kv[0])):
        s = s.replace(a, m)
    return s


def mask_command(cmd, task_id, sched_addr):
    parts = cmd.split(" ")
    out = []
    for p in parts:
        if p == sys.executable:
            out.append("<python>")
        elif p == task_id:
            out.append("<id>")
        elif p == sched_addr:
            out.append("<sched>")
        elif p.endswith(".server"):
            out.append("<pkg>.server")
        else:
            out.append(p)
    return " ".join(out)


def record_task_info(ti, sched_addr):
    res = []
    for r in ti.resources:
        if r.type == "SET":
            res.append([r.name, "SET", list(r.set.item)])
        else:
            res.append([r.name, "SCALAR", r.scalar.value])
    env = []
    for v in ti.command.environment.variables:
        env.append([v.name, MASKED_ENV_VALUES.get(v.name, v.value)])
    return {
        "name": ti.name,
        "agent": ti.agent_id.value,
        "resources": res,
        "shell": bool(ti.command.shell),
        "command": mask_command(ti.command.value, ti.task_id.value, sched_addr),
        "env": env,
    }


class RecordingDriver(object):
    def __init__(self, sched_addr):
        self.sched_addr = sched_addr
        self.events = []

    def launchTasks(self, offer_id, infos):
        self.events.append(["launch", offer_id.value,
                            [record_task_info(ti, self.sched_addr) for ti in infos]])

    def declineOffer(self, offer_id, filters=None):
        rs = None
        if filters is not None:
            rs = filters["refuse_seconds"]
        self.events.append(["decline", offer_id.value, rs])

    def suppressOffers(self):
        self.events.append(["suppress"])

    def reviveOffers(self):
        self.events.append(["revive"])


def task_table(sched):
    return [["%s:%s" % (t.job_name, t.task_index), bool(t.offered)]
            for t in sched.tasks.values()]


# ----------------------------------------------------------------------------
PLACEMENT_CASES = {
    # SURVEY.md 4.3 probe: ps:0->GPU-0 ... worker:3->GPU-5, second offer declined
    "ps2_w4_one_offer_8gpu": dict(
        jobs=[dict(name="ps", num=2, gpus=1), dict(name="worker", num=4, gpus=1)],
        rounds=[[dict(id="o1", cpus=64.0, mem=1e6,
                      gpus=["GPU-%d" % i for i in range(8)])],
                [dict(id="o2", cpus=64.0, mem=1e6, gpus=["GPU-8"])]]),
    # tfrun defaults: -Gs 0 (script/tfrun:25), -Gw 1
    "tfrun_s1_w2_gw1": dict(
        jobs=[dict(name="ps", num=1, cpus=1.0, gpus=0, mem=1024.0, cmd="x"),
              dict(name="worker", num=2, cpus=1.0, gpus=1, mem=1024.0, cmd="x")],
        rounds=[[dict(id="o1", cpus=8.0, mem=65536.0, gpus=["0", "1"])]]),
    "scalar_gpus": dict(
        jobs=[dict(name="worker", num=2, gpus=2)],
        rounds=[[dict(id="o1", cpus=8.0, mem=65536.0, gpus=4, gpu_type="SCALAR")]]),
    "gpu_starved_then_second_offer": dict(
        jobs=[dict(name="worker", num=3, gpus=1)],
        rounds=[[dict(id="o1", cpus=8.0, mem=65536.0, gpus=["a", "b"])],
                [dict(id="o2", cpus=8.0, mem=65536.0, gpus=["c", "d"])],
                [dict(id="o3", cpus=8.0, mem=65536.0, gpus=["e"])]]),
    "cpu_bound": dict(
        jobs=[dict(name="ps", num=1, cpus=1.0), dict(name="worker", num=3, cpus=1.0)],
        rounds=[[dict(id="o1", cpus=2.5, mem=65536.0)],
                [dict(id="o2", cpus=2.0, mem=65536.0)]]),
    "mem_bound": dict(
        jobs=[dict(name="worker", num=3, mem=4096.0)],
        rounds=[[dict(id="o1", cpus=8.0, mem=10000.0)]]),
    "fractional_gpu": dict(
        jobs=[dict(name="worker", num=2, gpus=0.5)],
        rounds=[[dict(id="o1", cpus=8.0, mem=65536.0, gpus=["g0", "g1", "g2"])]]),
    "job_start_offset": dict(
        jobs=[dict(name="ps", num=1), dict(name="worker", num=4, start=2, gpus=1)],
        rounds=[[dict(id="o1", cpus=8.0, mem=65536.0, gpus=["0", "1", "2", "3"])]]),
    "env_passthrough": dict(
        jobs=[dict(name="worker", num=1)],
        kw=dict(env={"FOO": "bar", "PYTHONPATH": "/dropped", "A": "1"}),
        rounds=[[dict(id="o1", cpus=8.0, mem=65536.0)]]),
    "no_gpu_resource_gpu_task": dict(
        jobs=[dict(name="ps", num=1), dict(name="worker", num=1, gpus=1)],
        rounds=[[dict(id="o1", cpus=8.0, mem=65536.0)]]),
    "two_offers_one_call": dict(
        jobs=[dict(name="worker", num=4, gpus=1)],
        rounds=[[dict(id="o1", cpus=8.0, mem=65536.0, gpus=["a", "b"], agent="A1"),
                 dict(id="o2", cpus=8.0, mem=65536.0, gpus=["c", "d", "e"], agent="A2")]]),
    "mixed_gpu_zero_and_more": dict(
        jobs=[dict(name="ps", num=2, gpus=0), dict(name="worker", num=6, gpus=1)],
        rounds=[[dict(id="o1", cpus=192.0, mem=2e6,
                      gpus=[str(i) for i in range(8)])]]),
    "big_gpu_task_skipped_small_fits": dict(
        jobs=[dict(name="a", num=1, gpus=4), dict(name="b", num=2, gpus=1)],
        rounds=[[dict(id="o1", cpus=8.0, mem=65536.0, gpus=["0", "1", "2"])]]),
}


def run_placement(make_scheduler, make_offer):
    out = {}
    for name, case in sorted(PLACEMENT_CASES.items()):
        sched = make_scheduler(case["jobs"], **case.get("kw", {}))
        sched.addr = "sched-host:4000"
        drv = RecordingDriver(sched.addr)
        sched.driver = drv
        rounds = []
        for offers in case["rounds"]:
            before = len(drv.events)
            sched.resourceOffers(drv, [make_offer(o) for o in offers])
            rounds.append({"events": drv.events[before:], "tasks": task_table(sched)})
        out[name] = rounds
        del sched.driver
    return out


# ----------------------------------------------------------------------------
CLUSTER_DEF_CASES = {
    "ps2_w2": dict(jobs=[dict(name="ps", num=2), dict(name="worker", num=2)]),
    "start_offset": dict(jobs=[dict(name="ps", num=1),
                               dict(name="worker", num=3, start=1, gpus=1, cpus=2.0)]),
    "replica_cmd": dict(jobs=[dict(name="ps", num=1, cmd="run {job_name}"),
                              dict(name="worker", num=2, cmd="run {job_name}", gpus=1)],
                        kw=dict(protocol="grpc+verbs", extra_config={"initializer": "true"},
                                forward_addresses={"/job:worker/task:0": ["h", 1]})),
}


def _fake_task(sched_addr, task_id, my_addr, sink):
    host, port = sched_addr.rsplit(":", 1)
    c = socket.create_connection((host, int(port)), timeout=30)
    c.sendall(_frame((task_id, my_addr)))
    cfg = _read_frame(c)
    c.sendall(_frame("ok"))
    c.close()
    sink[task_id] = cfg


def _frame(o):
    d = pickle.dumps(o)
    return struct.pack(">I", len(d)) + d


def _read_exact(c, n):
    buf = b""
    while len(buf) < n:
        chunk = c.recv(n - len(buf))
        if not chunk:
            raise EOFError("short read")
        buf += chunk
    return buf


def _read_frame(c):
    (n,) = struct.unpack(">I", _read_exact(c, 4))
    return pickle.loads(_read_exact(c, n))


class HandshakeDriver(object):
    """Fake driver: one generous offer; every launched TaskInfo becomes a thread
    that plays tfmesos/server.py:25-49 against the scheduler's rendez-vous socket."""
    version = "1.0.0"
    D = None

    def __init__(self, sched, framework=None, master=None, use_addict=False):
        self.sched = sched
        self.configs = {}
        self.threads = []
        self.infos = []

    def start(self):
        D = self.D
        fid, mi = D(), D()
        fid.value, mi.hostname, mi.port = "fw", "localhost", 5050
        self.sched.registered(self, fid, mi)
        offer = build_offer(D, dict(id="o1", cpus=512.0, mem=1e7,
                                    gpus=[str(i) for i in range(16)]))
        self.sched.resourceOffers(self, [offer])

    def launchTasks(self, offer_id, infos):
        for ti in infos:
            self.infos.append(ti)
            job, idx = ti.name[len("/job:"):].split("/task:")
            fake_addr = "%s-host-%s:%d" % (job, idx, 2000 + int(idx))
            t = threading.Thread(target=_fake_task, args=(
                self.sched.addr, ti.task_id.value, fake_addr, self.configs))
            t.daemon = True
            t.start()
            self.threads.append(t)

    def suppressOffers(self):
        pass

    def declineOffer(self, *a, **k):
        pass

    def reviveOffers(self):
        pass

    def stop(self):
        pass

    def join(self):
        pass


def run_cluster_def(make_scheduler, install_driver=None, D=None):
    """install_driver(cls) makes scheduler.start() construct `cls`; default
    patches the reference module global (scheduler.py:336)."""
    out = {}
    for name, case in sorted(CLUSTER_DEF_CASES.items()):
        sched = make_scheduler(case["jobs"], **case.get("kw", {}))
        holder = {}

        class Drv(HandshakeDriver):
            def __init__(self, *a, **k):
                HandshakeDriver.__init__(self, *a, **k)
                holder["drv"] = self

        if D is None:
            import addict
            Drv.D = addict.Dict
        else:
            Drv.D = D
        mod, orig_driver = None, None
        if install_driver is None:
            mod = sys.modules[type(sched).__module__]
            orig_driver = mod.MesosSchedulerDriver
            mod.MesosSchedulerDriver = Drv
        else:
            install_driver(sched, Drv)
        try:
            sched.start()
        finally:
            if mod is not None:
                mod.MesosSchedulerDriver = orig_driver
        drv = holder["drv"]
        for t in drv.threads:
            t.join(30)
        by_id = {t.mesos_task_id: t for t in sched.tasks.values()}
        import os
        configs = {}
        for tid, cfg in drv.configs.items():
            t = by_id[tid]
            cfg = dict(cfg)
            cfg["cwd"] = "<cwd>" if cfg["cwd"] == os.getcwd() else cfg["cwd"]
            configs["%s:%s" % (t.job_name, t.task_index)] = cfg
        out[name] = {"configs": configs, "targets": dict(sched.targets),
                     "started": bool(sched.started),
                     "initalized": [["%s:%s" % (t.job_name, t.task_index),
                                     bool(t.initalized)] for t in sched.tasks.values()]}
        sched.stop()
    return out


# ----------------------------------------------------------------------------
def _update(make_update, task_id, state, message="m"):
    u = make_update()
    u.task_id.value = task_id
    u.state = state
    u.message = message
    return u


def _ids(sched):
    return {"%s:%s" % (t.job_name, t.task_index): t.mesos_task_id
            for t in sched.tasks.values()}


def run_status(make_scheduler, make_update):
    out = {}
    jobs = [dict(name="ps", num=1), dict(name="worker", num=2)]

    # (a) pre-start failures: revive up to MAX_FAILURE_COUNT, then raise
    sched = make_scheduler(jobs)
    drv = RecordingDriver("x")
    sched.driver = drv
    log = []
    for attempt in range(4):
        ids = _ids(sched)
        old = ids["worker:0"]
        task = sched.tasks[old]
        task.offered = True
        try:
            sched.statusUpdate(drv, _update(make_update, old, "TASK_FAILED"))
            new = _ids(sched)["worker:0"]
            log.append({"raised": None, "new_id_differs": new != old,
                        "offered": bool(sched.tasks[new].offered),
                        "failure_count": dict(sched.task_failure_count),
                        "n_tasks": len(sched.tasks),
                        "order": [k for k in _ids(sched)],
                        "events": list(drv.events)})
        except RuntimeError:
            log.append({"raised": "RuntimeError",
                        "failure_count": dict(sched.task_failure_count)})
        drv.events = []
    out["prestart_failures"] = log
    del sched.driver

    # (b) post-start: FINISHED counts, finished() flips when ONE job is complete
    sched = make_scheduler(jobs)
    sched.started = True
    ids = _ids(sched)
    seq = []
    for key, state in [("worker:0", "TASK_RUNNING"), ("worker:0", "TASK_FINISHED"),
                       ("worker:1", "TASK_FINISHED")]:
        sched.statusUpdate(None, _update(make_update, ids[key], state))
        seq.append([key, state, dict(sched.job_finished), bool(sched.finished())])
    sched.statusUpdate(None, _update(make_update, "no-such-id", "TASK_FAILED"))
    seq.append(["unknown", "TASK_FAILED", dict(sched.job_finished), bool(sched.finished())])
    out["poststart_finish"] = seq

    # (c) post-start failure of any kind is fatal
    fatal = {}
    for state in ["TASK_FAILED", "TASK_KILLED", "TASK_ERROR", "TASK_LOST"]:
        sched = make_scheduler(jobs)
        sched.started = True
        try:
            sched.statusUpdate(None, _update(make_update, _ids(sched)["ps:0"], state))
            fatal[state] = None
        except RuntimeError:
            fatal[state] = "RuntimeError"
    out["poststart_fatal"] = fatal

    # (d) start>0 quirk: finished() compares with job.num (scheduler.py:474-477)
    sched = make_scheduler([dict(name="worker", num=3, start=1)])
    sched.started = True
    for tid in list(_ids(sched).values()):
        sched.statusUpdate(None, _update(make_update, tid, "TASK_FINISHED"))
    out["start_offset_finish"] = [dict(sched.job_finished), bool(sched.finished())]

    # (e) agent / executor loss and error()
    lost = {}
    for started in (False, True):
        sched = make_scheduler(jobs)
        sched.started = started
        row = {}
        for meth, args in [("slaveLost", (None, make_update(value="a1"))),
                           ("executorLost", (None, make_update(value="e1"),
                                             make_update(value="a1"), 1)),
                           ("error", (None, "boom"))]:
            try:
                getattr(sched, meth)(*args)
                row[meth] = None
            except RuntimeError:
                row[meth] = "RuntimeError"
        lost[str(started)] = row
    out["loss"] = lost
    return out


# ----------------------------------------------------------------------------
def run_job_normalisation(cluster, Job, sched_module, attr="TFMesosScheduler",
                          holder_module=None):
    """tfmesos/__init__.py:7-22: dict | Job | list of dict/Job -> [Job]; start()
    then always stop()."""
    holder = holder_module if holder_module is not None else sys.modules[cluster.__module__]
    seen = []

    class Recorder(object):
        def __init__(self, jobs, **kw):
            self.jobs = jobs
            self.kw = kw
            self.calls = []
            seen.append(self)

        def start(self):
            self.calls.append("start")

        def stop(self):
            self.calls.append("stop")

    orig = getattr(holder, attr)
    setattr(holder, attr, Recorder)
    out = {}
    try:
        forms = {
            "dict": dict(name="worker", num=2),
            "job": Job("ps", 1, gpus=1),
            "list_mixed": [dict(name="ps", num=1), Job("worker", 3, cpus=2.0, start=1)],
        }
        for k, v in sorted(forms.items()):
            with cluster(v, master="m", quiet=True) as s:
                s.calls.append("body")
            rec = seen[-1]
            out[k] = {"jobs": [[j.name, j.num, j.cpus, j.mem, j.gpus, j.cmd, j.start]
                               for j in rec.jobs],
                      "all_jobs": all(isinstance(j, Job) for j in rec.jobs),
                      "kw": sorted(rec.kw.items()), "calls": rec.calls}
        try:
            with cluster(dict(name="w", num=1), master="m") as s:
                raise ValueError("body failed")
        except ValueError:
            out["body_raises"] = seen[-1].calls
    finally:
        setattr(holder, attr, orig)
    return out


# ----------------------------------------------------------------------------
def run_wire(utils):
    """utils.send/recv framing (tfmesos/utils.py:6-15) over a socketpair."""
    a, b = socket.socketpair()
    objs = [("id-1", "host:1234"), "ok",
            {"job_name": "ps", "task_index": 0, "cluster_def": {"ps": ["a:1"]}}]
    out = []
    for o in objs:
        utils.send(a, o)
        raw = _read_exact(b, 4)
        (n,) = struct.unpack(">I", raw)
        payload = _read_exact(b, n)
        out.append({"len_prefix_ok": n == len(payload),
                    "roundtrip": pickle.loads(payload) == o})
        b.sendall(_frame(o))
        back = utils.recv(a)
        out[-1]["recv_ok"] = back == o or list(back) == list(o)
    a.close()
    b.close()
    return out
