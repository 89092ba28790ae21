"""A/B timing of kernel variants: python profiles/scripts/ab_phase.py lib1.so lib2.so ...  (run on the GPU box)
Each library (an `ab` build of csrc with -D switches) aligns the same resident batch (BASELINE configs[2] shape unless
AB_PAIRS / AB_N / AB_E / AB_PRESET say otherwise) in its own process; prints the CUDA-event time of each phase kernel
(median of AB_REPS runs after one warm-up) and checks that the costs and CIGAR digests equal those of the first library."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

WORKER_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'resident_run.py')


def main():
    ref = None
    for lib in sys.argv[1:]:
        env = dict(os.environ, APA_LIB=os.path.abspath(lib))
        out = subprocess.run([sys.executable, WORKER_PATH], env=env, capture_output=True, text=True)
        if out.returncode != 0:
            print(lib, "FAILED", out.stderr[-400:])
            continue
        r = json.loads(out.stdout.strip().splitlines()[-1])
        ref = ref if ref is not None else r["digest"]
        print("%-28s build %6.2f  pass %6.2f  trace %6.2f  total %7.2f  retries %d  %s" % (
            os.path.basename(lib), *r["phase_ms"], r["kernel_ms"], r["retries"], "same results" if r["digest"] == ref else "RESULTS DIFFER"))


if __name__ == "__main__":
    main()
