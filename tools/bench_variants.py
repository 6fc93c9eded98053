"""Times experimental builds of libntm_mpc.so side by side on the headline workload (config 3, fixed(10), resident
inputs, CUDA events).  Build them with tools/build_variants.sh into mpc-ntm-control_b200/lib/variants/; each one runs
in its own process (NTM_MPC_LIB) and prints the median of 5 launches plus a checksum of uk."""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "--one":
    sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "mpc-ntm-control_b200"))
    import numpy as np, torch
    import ntm_mpc
    from ntm_mpc import physics
    mpc = ntm_mpc.NtmMpc(0); dev = torch.device("cuda:0")
    mpc.set_stream(torch.cuda.current_stream().cuda_stream)
    for cfg, S, prof in ((3, 65536, 16), (3, 65536, 0), (5, 1024, 16)):
        P, x0, N = physics.batch_params(cfg, S=S)
        dx = torch.from_numpy(x0).to(dev); dP = torch.from_numpy(np.ascontiguousarray(P.T)).to(dev)
        xk = torch.empty((S, 21, 2), dtype=torch.float64, device=dev); uk = torch.empty((S, 20), dtype=torch.float64, device=dev)
        st = torch.empty((S,), dtype=torch.int32, device=dev)
        run = lambda: mpc.closed_loop_dev(S, N, 20, 10, 1e-14, prof, 0, dx.data_ptr(), dP.data_ptr(), S, xk.data_ptr(), uk.data_ptr(), 0, 0, 0, 0, st.data_ptr())
        run(); torch.cuda.synchronize(); ts = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); run(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
        print(f"   config{cfg} prof {prof:2d}: {sorted(ts)[2]:8.3f} ms  checksum {uk.sum().item():.10e} status {int(st.max())}", flush=True)
    sys.exit(0)
libs = [os.path.join(ROOT, "mpc-ntm-control_b200", "lib", "libntm_mpc.so")] + sorted(glob.glob(os.path.join(ROOT, "mpc-ntm-control_b200", "lib", "variants", "*.so")))
for lib in libs:
    print(os.path.relpath(lib, ROOT), flush=True)
    subprocess.run([sys.executable, os.path.abspath(__file__), "--one"], env=dict(os.environ, NTM_MPC_LIB=lib))
