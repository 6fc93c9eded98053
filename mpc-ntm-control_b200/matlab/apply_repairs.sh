#!/bin/sh
# Produces NTM_MPC_Sim_fixed.m from an unmodified checkout of IsaacSavona/MPC-NTM-Control: the five one-line repairs of
# INTEGRATION.md section 4 (SURVEY defects D1, D8, D9, D12 and the toolbox call), addressed by line number so that no
# reference text has to live in this repository.  Usage:  sh apply_repairs.sh /path/to/MPC-NTM-Control
# The result calls the MEX shims built per INTEGRATION.md section 2 (rho1, rho2, rho3, A, B, Rho_to_PhiGammaLambda,
# ntm_qp_box).  Untested here: the build image has neither MATLAB nor Octave.
set -e
src="$1/NTM_MPC_Sim.m"
[ -f "$src" ] || { echo "usage: sh apply_repairs.sh <reference checkout>" >&2; exit 2; }
[ "$(wc -l < "$src")" -eq 164 ] || { echo "unexpected NTM_MPC_Sim.m (not 164 lines): refusing to patch by line number" >&2; exit 3; }
sed -e '54d' -e '56d' \
    -e '71s/.*/R = repmat(r(:),N,1); % compact notation of R (2N x 1)/' \
    -e '74s/.*/% state-constraint rows dropped: EC-power box only (see getWLc MEX for W, L, c)/' \
    -e '88s/.*/% quadprog options not needed: ntm_qp_box/' \
    -e '97s/.*/            [U,exitflag] = ntm_qp_box(G,F,umin,umax); % EC-power-bounded QP on the GPU/' \
    -e '130s/.*/    xcur = xk(:,k); xk(:,k+1) = A(rho1(xcur), rho2(xcur))*xcur + B(rho3(xcur))*uk(:,k); % evolve state one time step/' \
    "$src" > NTM_MPC_Sim_fixed.m
echo "wrote NTM_MPC_Sim_fixed.m"
