function P = ntm_params(c)
%NTM_PARAMS  16 x S parameter block of include/ntm_mpc.h from a struct of physics constants.
%   c has the fields the reference script defines (NTM_MPC_Sim.m:5-25,31,47-48,59-60): j_BS, w_dep, w_marg, w_sat,
%   tau_r, rs, a, eta_CD, tau_E0, mu0, Lq, B_pol, m, Cw, tau_A0, tau_w, omega0, Ts, umin, umax, r (1x2), Q (2x2).
%   Every field may be a scalar or a 1 x S row (one column of P per scenario).
%   The products are hoisted in the evaluation order of A.m:2 / B.m:2 so the GPU rounds like the .m files.
kappa = 16*c.mu0.*c.Lq.*c.rs.^2./(0.82*c.tau_r.*c.B_pol*pi);            % NTM_MPC_Sim.m:24
zeta  = c.m.*c.Cw.*c.tau_A0.^2.*c.tau_w.*c.a.^3;                        % :25
C1 = -4/3*(kappa.*c.Ts.*c.j_BS.*c.w_sat)./(c.w_sat.^2 + c.w_marg.^2);   % :37
C2 = c.Ts.*c.omega0./c.tau_E0;
tau_E = c.tau_E0;                                                        % :14
rows = {(4/3)*(kappa.*c.rs./(0.82*c.tau_r)).*c.Ts, c.Ts./(zeta.*c.a.^3), 1 - c.Ts./tau_E, ...
        (kappa.*c.Ts.*c.eta_CD./c.w_dep), C1, C2, c.w_marg.^2, c.w_dep, c.umin, c.umax, ...
        c.r(1), c.r(2), c.Q(1,1), c.Q(1,2), c.Q(2,2), 0};
S = max(cellfun(@numel, rows));
P = zeros(16, S);
for i = 1:16
    P(i,:) = rows{i}(:).' .* ones(1, S);
end
end
