"""Expected DRAM bytes per site of the full-lattice gather kernels (link construction, fermion force) in the two launch
orders, next to what ncu measured (profiles/ncu_staple_r02u.txt, profiles/ncu_fbwd_r02v.txt).

Model: a matrix field is 9 planes of double2 = 144 B per site, stored parity-blocked (all even sites, then all odd).
A neighbour at an odd number of hops has the other parity.  In the plain order (thread i = site i) the first half of
the grid works on even sites, the second on odd ones, a gigabyte of traffic apart, so nothing survives in the 126 MB L2
between them: a field read at both parities (relative to the site) is streamed once per half of the grid = 2 x 144 B per
site; a field read at one parity only is streamed once.  With the parities interleaved CTA by CTA every field is
streamed once.  Outputs: a store is 144 B, a read-modify-write 288 B.

    python profiles/traffic_model.py
"""
M = 144  # bytes of one 3x3 complex double matrix

# loads as (field, hops from the site): parity relative to the site = hops % 2
KERNELS = {
    # links.cuh staple_kernel: upper U(x) L(x+nu) U(x+mu)^+, lower U(x-nu)^+ L(x-nu) U(x-nu+mu)
    "staple_kernel<true>  (stores the staple, fat RMW)": dict(
        loads=[("U", 0), ("L", 1), ("U", 1), ("U", 1), ("L", 1), ("U", 2)], stores=1, rmw=1, measured=(867, 709)),
    "staple_kernel<false> (fat RMW only)": dict(
        loads=[("U", 0), ("L", 1), ("U", 1), ("U", 1), ("L", 1), ("U", 2)], stores=0, rmw=1, measured=(713, 566)),
    # force.cuh StapleBwdSite, part = 3: U at z+mu, z-mu, z-nu, z-nu+mu, z; H at z-nu, z, z-mu, z+nu, z-mu+nu;
    # L at z+nu, z-mu+nu, z-mu, z
    "StapleBwdSite, full pass (two gradient RMWs)": dict(
        loads=[("U", 1), ("U", 1), ("U", 1), ("U", 2), ("U", 0), ("H", 1), ("H", 0), ("H", 1), ("H", 1), ("H", 2),
               ("L", 1), ("L", 2), ("L", 1), ("L", 0)], stores=0, rmw=2, measured=(1451, 1009)),
    # part = 1 (upper staple): U at z-nu, z-nu+mu, z+mu, z-mu; H at z-nu, z, z-mu; L at z+nu, z-mu+nu
    "StapleBwdSite, half pass (upper staple)": dict(
        loads=[("U", 1), ("U", 2), ("U", 1), ("U", 1), ("H", 1), ("H", 0), ("H", 1), ("L", 1), ("L", 2)], stores=0, rmw=2,
        measured=(1432, 1001)),
    # force.cuh StapleFwdSite: as staple_kernel without the fat link
    "StapleFwdSite (stores the staple)": dict(
        loads=[("U", 0), ("L", 1), ("U", 1), ("U", 1), ("L", 1), ("U", 2)], stores=1, rmw=0, measured=(None, None)),
}


def expected(k, interleaved):
    par = {}
    for f, hops in k["loads"]:
        par.setdefault(f, set()).add(hops % 2)
    reads = sum(M * (1 if interleaved else len(p)) for p in par.values())
    return reads + M * k["stores"] + 2 * M * k["rmw"]


def main():
    print("%-52s %28s %28s" % ("kernel", "even sites, then odd", "parities interleaved"))
    print("%-52s %14s %13s %14s %13s" % ("", "expected", "measured", "expected", "measured"))
    for name, k in KERNELS.items():
        e0, e1 = expected(k, False), expected(k, True)
        m0, m1 = k["measured"]
        print("%-52s %10d B/site %9s %10d B/site %9s" % (name, e0, "-" if m0 is None else "%d" % m0, e1, "-" if m1 is None else "%d" % m1))
    print("(measured: averages of dram__bytes_read.sum + dram__bytes_write.sum per launch / 2097152 sites, 32^3 x 64;")
    print(" the lower-staple half pass touches every field at both parities as well: same expectation)")


if __name__ == "__main__":
    main()
