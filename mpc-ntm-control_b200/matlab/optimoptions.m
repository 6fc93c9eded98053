function opt = optimoptions(solver, varargin)
%OPTIMOPTIONS  Stand-in for installations without the Optimization Toolbox.
%   NTM_MPC_Sim.m:88 builds  opt = optimoptions('quadprog','Display','off')  and passes it as the tenth
%   argument of quadprog (:97).  The quadprog MEX shim (csrc/mex/ntm_mex.c, gateway 10) ignores that
%   argument, so a plain struct carrying the name/value pairs is all that is needed.  Delete this file
%   when the real toolbox is on the path.
opt = struct('SolverName', solver);
for k = 1:2:numel(varargin) - 1
    opt.(varargin{k}) = varargin{k + 1};
end
end
