%NTM_MPC_DEMO  The closed loop of NTM_MPC_Sim.m:93-131 for a batch of scenarios through the B200 MEX entry.
%   Build the MEX files first (INTEGRATION.md).  Untested here: the build image has neither MATLAB nor Octave.
c = struct('j_BS',73e3,'w_dep',0.024,'w_marg',0.02,'w_sat',0.32,'tau_r',293,'rs',1.55,'a',2.0,'eta_CD',0.9, ...
           'tau_E0',3.7,'mu0',4e-7*pi,'Lq',0.87,'B_pol',0.97,'m',2,'Cw',1,'tau_A0',3e-6,'tau_w',0.188, ...
           'omega0',2*pi*420,'Ts',0.1,'umin',0,'umax',2e6,'r',[0 1000*2*pi],'Q',eye(2));
S = 1024; N = 10; k_sim = 20; i_sim = 10; epsilon = 1e-14;
x0 = [0.06 + 0.09*rand(1,S); 1000*2*pi*ones(1,S)];        % BASELINE config 2
profile = 0;                                               % 0 = literal reading; +16 = always i_sim inner iterations
[xk, uk, cost, inner, status] = ntm_mpc_batch(x0, ntm_params(c), N, k_sim, i_sim, epsilon, profile);
w = xk(1:2:end, :);                                        % island width trajectories, (k_sim+1) x S
stairs(0:k_sim-1, uk(:,1)); xlabel('k'); ylabel('P_{ECCD} [W]');
