"""Run every GPU check group in its own process (a CUDA fault in one group cannot poison the others)
and print one table.  Usage on the B200 box:

    python tests/run_gpu_diag.py [group ...] > gpurun_out/diag.log
"""
import os
import subprocess
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))


def run_group(name):
    sys.path.insert(0, HERE)
    import torch
    import gpu_checks as G
    golden = G.load_golden()
    t0 = time.time()
    try:
        res = G.all_groups()[name](golden)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        import traceback
        traceback.print_exc()
        print(f"GROUP {name}: EXCEPTION {type(e).__name__}: {str(e)[:400]}")
        return 1
    bad = 0
    for label, err, tol in res:
        ok = (err <= tol) and err == err
        bad += not ok
        print(f"  {'ok  ' if ok else 'FAIL'} {label:<70s} err={err:.3e} tol={tol:.1e}")
    print(f"GROUP {name}: {len(res) - bad}/{len(res)} ok in {time.time() - t0:.1f}s")
    return 1 if bad else 0


def main():
    if len(sys.argv) > 2 and sys.argv[1] == "--one":
        sys.exit(run_group(sys.argv[2]))
    sys.path.insert(0, HERE)
    import gpu_checks as G
    names = sys.argv[1:] or [n for n in G.all_groups() if not n.startswith('calib')]
    failed = []
    for n in names:
        print(f"===== {n} =====", flush=True)
        try:
            p = subprocess.run([sys.executable, os.path.abspath(__file__), "--one", n], timeout=600)
            if p.returncode:
                failed.append(n)
        except subprocess.TimeoutExpired:
            print(f"GROUP {n}: TIMEOUT")
            failed.append(n)
        sys.stdout.flush()
    print("FAILED GROUPS:", failed if failed else "none")
    sys.exit(1 if failed else 0)


if __name__ == "__main__":
    main()
