# JuqboxB200.jl — the reference-side binding of libjuqbox_b200.so (include/juqbox_b200.h).
#
# NOT EXECUTED IN THIS REPO'S CI: the build image has no `julia`.  It is the ~150 lines a Juqbox maintainer adds
# (e.g. as src/b200.jl, `include`d from src/Juqbox.jl after evalobjgrad.jl and ipopt_interface.jl) so that
# setup scripts, `setup_ipopt_problem` and `run_optimizer` run unchanged while every objective/gradient
# evaluation goes to the GPU.  The Python mirror juqbox_b200/api.py is this file's tested twin.
#
# What it replaces in the reference:
#   Working_Arrays(params, nCoeff)                      src/evalobjgrad.jl:405      -> Working_Arrays_B200
#   traceobjgrad(pcof0, params, wa, verbose, evaladjoint)  src/evalobjgrad.jl:504   -> new method on Working_Arrays_B200
#   eval_f_g_grad!(pcof, params, wa, nodes, weights, ..)   src/ipopt_interface.jl:24 -> new method: ONE batched ccall
#                                                                                     instead of the nquad loop
# eval_f_par / eval_grad_f_par / eval_g_par / eval_jac_g_par take `wa` untyped (src/ipopt_interface.jl:77-179) and
# need no change: they dispatch to the methods below through `wa`.

const libjq = get(ENV, "JUQBOX_B200_LIB", "libjuqbox_b200.so")

struct JqOperator            # == jq_operator
    format::Int32
    nnz::Int64
    colptr::Ptr{Int64}
    rowval::Ptr{Int64}
    nzval::Ptr{Float64}
end

struct JqProblem             # == jq_problem
    n::Int32; m::Int32; ncoupled::Int32; nfreq::Int32
    neumann_terms::Int32; obj_func_type::Int32; pfid_type::Int32; linear_solver::Int32
    nsteps::Int64
    T::Float64
    uinit::Ptr{Float64}; vtarget_r::Ptr{Float64}; vtarget_i::Ptr{Float64}; wdiag::Ptr{Float64}; cfreq::Ptr{Float64}
    h0::JqOperator
    hsym::Ptr{JqOperator}
    hanti::Ptr{JqOperator}
    solver_tol::Float64
end

jq_error() = unsafe_string(ccall((:jq_last_error, libjq), Cstring, ()))
jq_check(rc) = rc == 0 ? nothing : error("juqbox_b200: ", jq_error())     # non-zero -> Julia error(), as the reference does

mutable struct Working_Arrays_B200
    handle::Ptr{Cvoid}
    nCoeff::Int64
    keep::Vector{Any}        # arrays referenced by the descriptor while jq_create runs
end

# dense Array{Float64,2} is already column-major; SparseMatrixCSC is CSC with 1-based indices -> shift to 0-based
function jq_operator(A::Array{Float64,2}, keep)
    push!(keep, A)
    JqOperator(0, length(A), C_NULL, C_NULL, pointer(A))
end
function jq_operator(A::SparseMatrixCSC{Float64,Int64}, keep)
    cp = A.colptr .- 1; rv = A.rowval .- 1; nz = copy(A.nzval)
    append!(keep, (cp, rv, nz))
    JqOperator(1, length(nz), pointer(cp), pointer(rv), pointer(nz))
end

function Working_Arrays_B200(params::objparams, nCoeff::Int64; device::Int = 0)
    @assert params.Nunc == 0 "uncoupled controls stay on the CPU path"
    @assert params.linear_solver.solver_id in (NEUMANN_SOLVER, JACOBI_SOLVER) "Neumann and Jacobi solvers are built for B200"
    @assert params.Integrator_id == Stormer_Verlet
    @assert isa(params.wmat_real, Diagonal) "custom forbidden-state weights stay on the CPU path"
    keep = Any[]
    Ntot = params.N + params.Nguard
    Cf = Array{Float64,2}(params.Cfreq[1:params.Ncoupled, :])
    wd = Vector{Float64}(diag(params.wmat_real))
    hs = [jq_operator(h, keep) for h in params.Hsym_ops]
    ha = [jq_operator(h, keep) for h in params.Hanti_ops]
    append!(keep, (Cf, wd, hs, ha, params.Uinit, params.Utarget_r, params.Utarget_i))
    pb = JqProblem(Ntot, params.N, params.Ncoupled, params.Nfreq, params.linear_solver.max_iter, params.objFuncType,
                   params.pFidType, params.linear_solver.solver_id, params.nsteps, params.T, pointer(params.Uinit), pointer(params.Utarget_r),
                   pointer(params.Utarget_i), pointer(wd), pointer(Cf), jq_operator(params.Hconst, keep), pointer(hs), pointer(ha),
                   params.linear_solver.tol)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    GC.@preserve keep pb jq_check(ccall((:jq_create, libjq), Cint, (Ref{JqProblem}, Cint, Ref{Ptr{Cvoid}}), pb, device, h))
    wa = Working_Arrays_B200(h[], nCoeff, Any[])
    finalizer(w -> ccall((:jq_destroy, libjq), Cint, (Ptr{Cvoid},), w.handle), wa)
    return wa
end

# change_target! (src/evalobjgrad.jl:1492) must be followed by this to refresh the device copy
update_target!(wa::Working_Arrays_B200, params::objparams) =
    jq_check(ccall((:jq_update_target, libjq), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), wa.handle, params.Utarget_r, params.Utarget_i))

# Batched core: pcof is Npar x nbatch (one candidate per column), shifts is n x nsamples (one diagonal per column).
function traceobjgrad_batch(pcof::Array{Float64,2}, params::objparams, wa::Working_Arrays_B200;
                            shifts::Union{Nothing,Array{Float64,2}} = nothing, weights::Union{Nothing,Vector{Float64}} = nothing,
                            evaladjoint::Bool = true)
    Npar, nbatch = size(pcof)
    nsamples = shifts === nothing ? 1 : size(shifts, 2)
    nout = weights === nothing ? nbatch * nsamples : nbatch
    infid = zeros(nout); leak = zeros(nout); tinf = zeros(nout)
    two = params.objFuncType != 1                        # objFuncType 1: infidelgrad aliases totalgrad (src/evalobjgrad.jl:951),
    grad = zeros(Npar, nout)                             # so only one gradient is fetched from the device
    igrad = two ? zeros(Npar, nout) : grad
    lgrad = two ? zeros(Npar, nout) : zeros(Npar, 0)
    jq_check(ccall((:jq_traceobjgrad_batch, libjq), Cint,
                   (Ptr{Cvoid}, Int32, Ptr{Float64}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Int32,
                    Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                   wa.handle, nbatch, pcof, Npar, nsamples, shifts === nothing ? C_NULL : shifts,
                   weights === nothing ? C_NULL : weights, evaladjoint, infid, leak, tinf, grad,
                   two ? pointer(igrad) : Ptr{Float64}(C_NULL), two ? pointer(lgrad) : Ptr{Float64}(C_NULL)))
    return infid, leak, tinf, grad, igrad, lgrad
end

# Drop-in method: same return tuples as src/evalobjgrad.jl:1032-1035.  verbose=true is not offered on the GPU
# (state history + forward-sensitivity check, SURVEY.md row 14): keep a CPU Working_Arrays for plot_results.
function traceobjgrad(pcof0::Array{Float64,1}, params::objparams, wa::Working_Arrays_B200, verbose::Bool = false, evaladjoint::Bool = true)
    verbose && error("verbose=true: call traceobjgrad with a CPU Working_Arrays")
    infid, leak, tinf, grad, igrad, lgrad = traceobjgrad_batch(reshape(pcof0, :, 1), params, wa; evaladjoint = evaladjoint)
    objfv = infid[1] + leak[1]
    evaladjoint || return objfv, infid[1], leak[1]
    totalgrad = grad[:, 1]
    if params.objFuncType != 1
        return objfv, totalgrad, infid[1], leak[1], tinf[1], igrad[:, 1], lgrad[:, 1]
    else
        return objfv, totalgrad, infid[1], leak[1], tinf[1], totalgrad, zeros(0)
    end
end

# The risk-neutral sample loop (src/ipopt_interface.jl:38-65) as one call: no mutate/restore of params.Hconst.
function eval_f_g_grad!(pcof::Vector{Float64}, params::objparams, wa::Working_Arrays_B200,
                        nodes::AbstractArray = [0.0], weights::AbstractArray = [1.0], compute_adjoint::Bool = true)
    n = params.N + params.Nguard
    shifts = zeros(n, length(nodes))
    for i in 1:length(nodes), j in 2:n
        shifts[j, i] = 0.01 * nodes[i] * (10.0^(j - 2))          # src/ipopt_interface.jl:43
    end
    infid, leak, _, _, igrad, lgrad = traceobjgrad_batch(reshape(pcof, :, 1), params, wa; shifts = shifts,
                                                         weights = Vector{Float64}(weights), evaladjoint = compute_adjoint)
    params.last_pcof .= pcof
    params.last_infidelity = infid[1]
    params.last_leak = leak[1]
    if compute_adjoint
        params.last_infidelity_grad .= igrad[:, 1]                 # == total gradient when objFuncType == 1 (:951)
        params.objFuncType != 1 && (params.last_leak_grad .= lgrad[:, 1])
    end
    params.lastTraceInfidelity = params.last_infidelity
    params.lastLeakIntegral = params.last_leak
end

# evalctrl (src/plotstatectrl.jl:246-276) on the GPU handle: control `jFunc` (1-based) on the grid `td`, rad/ns.
function evalctrl(params::objparams, pcof0::Array{Float64,1}, td::Array{Float64,1}, jFunc::Int64, wa::Working_Arrays_B200)
    nt = length(td)
    p = zeros(nt, params.Ncoupled); q = zeros(nt, params.Ncoupled)       # column c = control c  ==  ABI [ncoupled][ntimes]
    jq_check(ccall((:jq_eval_controls, libjq), Cint, (Ptr{Cvoid}, Ptr{Float64}, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                   wa.handle, pcof0, length(pcof0), nt, td, p, q))
    return p[:, jFunc], q[:, jFunc]
end

# eval_forward(U0, pcof, params; saveEndOnly=false, saveEvery) (src/evalobjgrad.jl:2727-2873) with U0 = params.Uinit:
# returns the Ntot x N x nsave complex state history.
function eval_forward(pcof0::Array{Float64,1}, params::objparams, wa::Working_Arrays_B200; saveEvery::Int64 = 1)
    Ntot = params.N + params.Nguard
    nsave = div(params.nsteps, saveEvery) + 1
    hr = zeros(Ntot, params.N, nsave); hi = zeros(Ntot, params.N, nsave)
    jq_check(ccall((:jq_eval_forward, libjq), Cint,
                   (Ptr{Cvoid}, Int32, Ptr{Float64}, Int32, Int32, Ptr{Float64}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}),
                   wa.handle, 1, pcof0, length(pcof0), 1, C_NULL, saveEvery, hr, hi, C_NULL, C_NULL))
    return hr .+ im .* hi
end
