"""Child process of bench.py: samples SM clock + throttle reasons of one GPU through NVML as fast as NVML answers, with wall
timestamps, until its stdin closes; then prints the samples as one JSON line.  A separate PROCESS (not a thread) so that the
sampler never competes with the launch loop for the interpreter lock: a 20-step timed region lasts half a millisecond.

    python benchmarks/_clock_sampler.py <nvml device index>
"""
import json
import select
import sys
import time


def main() -> None:
    index = int(sys.argv[1])
    out = {"max_mhz": None, "samples": [], "error": None}
    try:
        import pynvml as nv

        nv.nvmlInit()
        h = nv.nvmlDeviceGetHandleByIndex(index)
        out["max_mhz"] = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
        try:
            nv.nvmlDeviceGetCurrentClocksEventReasons(h)
            reasons = nv.nvmlDeviceGetCurrentClocksEventReasons
        except Exception:
            reasons = nv.nvmlDeviceGetCurrentClocksThrottleReasons
    except Exception as err:  # no NVML: report nothing rather than guess
        out["error"] = repr(err)
        nv = None
    sys.stdout.write("ready\n")
    sys.stdout.flush()
    samples = out["samples"]
    while True:
        if select.select([sys.stdin], [], [], 0.0001)[0]:  # parent closed the pipe (or wrote): stop
            break
        if nv is None:
            time.sleep(0.001)
            continue
        try:
            t = time.time()
            samples.append((t, nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), int(reasons(h))))
        except Exception:
            pass
        if len(samples) > 2_000_000:
            break
    sys.stdout.write(json.dumps(out) + "\n")
    sys.stdout.flush()


if __name__ == "__main__":
    main()
