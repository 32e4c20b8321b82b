"""Child command used by the replica-mode golden run: records what a task
launched through tfmesos/server.py:67-113 actually sees (expanded argv,
TFMESOS_* env, cwd)."""
import json
import os
import sys


def main():
    out_dir = sys.argv[1]
    env = {k: v for k, v in os.environ.items()
           if k.startswith("TFMESOS_") or k == "PYTHONUNBUFFERED"}
    rec = {"argv": sys.argv[2:], "env": env, "cwd": os.getcwd()}
    name = "%s_%s.json" % (env.get("TFMESOS_JOB_NAME"), env.get("TFMESOS_TASK_INDEX"))
    with open(os.path.join(out_dir, name), "w") as f:
        json.dump(rec, f)
    print("probe child", name)


if __name__ == "__main__":
    main()
