#!/bin/sh
# Produces NTM_MPC_Sim_fixed.m from an unmodified checkout of IsaacSavona/MPC-NTM-Control: the one-line repairs of
# INTEGRATION.md section 4 (SURVEY defects D1, D8, D9, D12 and the toolbox call), addressed by line number so that no
# reference text has to live in this repository.
#   sh apply_repairs.sh /path/to/MPC-NTM-Control              EC-power box only: line 97 calls ntm_qp_box
#   sh apply_repairs.sh /path/to/MPC-NTM-Control state-rows   state rows kept: line 74 calls getWLc with its own 7-argument
#                                                             signature, lines 88 and 97 stay as written and resolve to
#                                                             the optimoptions.m stub of this folder and the quadprog MEX
# The result calls the MEX shims built per INTEGRATION.md section 2.  Untested here: the build image has neither MATLAB
# nor Octave.  Note for the state-rows mode: the script's own x0 violates its own min_width (defect D18), so every QP
# reports exitflag -2 until x0(1) >= min_width.
set -e
src="$1/NTM_MPC_Sim.m"
mode="${2:-box}"
[ -f "$src" ] || { echo "usage: sh apply_repairs.sh <reference checkout> [state-rows]" >&2; exit 2; }
[ "$(wc -l < "$src")" -eq 164 ] || { echo "unexpected NTM_MPC_Sim.m (not 164 lines): refusing to patch by line number" >&2; exit 3; }
if [ "$mode" = "state-rows" ]; then
sed -e '54d' -e '56d' \
    -e '71s/.*/R = repmat(r(:),N,1); % compact notation of R (2N x 1)/' \
    -e '74s/.*/[W, L, c] = getWLc(xmax, xmin, umax, umin, Gamma, Phi, Lambda); % the 7-argument signature of getWLc.m:1/' \
    -e '130s/.*/    xcur = xk(:,k); xk(:,k+1) = A(rho1(xcur), rho2(xcur))*xcur + B(rho3(xcur))*uk(:,k); % evolve state one time step/' \
    "$src" > NTM_MPC_Sim_fixed.m
else
sed -e '54d' -e '56d' \
    -e '71s/.*/R = repmat(r(:),N,1); % compact notation of R (2N x 1)/' \
    -e '74s/.*/% state-constraint rows dropped: EC-power box only (see getWLc MEX for W, L, c)/' \
    -e '88s/.*/% quadprog options not needed: ntm_qp_box/' \
    -e '97s/.*/            [U,exitflag] = ntm_qp_box(G,F,umin,umax); % EC-power-bounded QP on the GPU/' \
    -e '130s/.*/    xcur = xk(:,k); xk(:,k+1) = A(rho1(xcur), rho2(xcur))*xcur + B(rho3(xcur))*uk(:,k); % evolve state one time step/' \
    "$src" > NTM_MPC_Sim_fixed.m
fi
echo "wrote NTM_MPC_Sim_fixed.m ($mode)"
