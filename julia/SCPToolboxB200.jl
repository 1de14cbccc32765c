#= SCPToolboxB200.jl -- thin ccall glue over libscpb (include/scpb.h).

NOT EXECUTABLE IN THE BUILD IMAGE (no Julia).  Kept mechanically obvious: one wrapper per C entry point, no logic.
See INTEGRATION.md for where each call replaces reference code. =#
module SCPToolboxB200

const libscpb = joinpath(@__DIR__, "..", "scptoolbox.jl_b200", "libscpb.so")

mutable struct Handle
    ptr::Ptr{Cvoid}
end

function Handle(device::Integer = 0)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    rc = ccall((:scpb_create, libscpb), Int32, (Int32, Ptr{Ptr{Cvoid}}), device, out)
    rc == 0 || error("scpb_create failed ($rc): no usable CUDA device")
    h = Handle(out[])
    finalizer(x -> ccall((:scpb_destroy, libscpb), Int32, (Ptr{Cvoid},), x.ptr), h)
    return h
end

function last_error(h::Handle)
    buf = Vector{UInt8}(undef, 512)
    ccall((:scpb_last_error, libscpb), Int32, (Ptr{Cvoid}, Ptr{UInt8}, Csize_t), h.ptr, buf, 512)
    return unsafe_string(pointer(buf))
end

check(h::Handle, rc::Int32, what) = rc == 0 || error("$what failed ($rc): $(last_error(h))")

model_set!(h::Handle, id::Integer, par::Vector{Float64}, nx, nu, np) = check(h,
    ccall((:scpb_model_set, libscpb), Int32, (Ptr{Cvoid}, Int32, Ptr{Float64}, Int32, Int32, Int32, Int32),
          h.ptr, id, par, length(par), nx, nu, np), "scpb_model_set")

""" discretize!: arrays are Julia column-major `xd[nx,N,B]`, `A[nx,nx,N-1,B]`, ... (B = 1 ⇒ the fields of `ref`). """
function discretize!(h::Handle, t_grid, xd, ud, p, iSx, feas_tol, Nsub, A, Bm, Bp, F, r, E, defect, feas::Vector{Int32})
    N = length(t_grid); B = length(feas); secs = Ref{Float64}(0.0)
    check(h, ccall((:scpb_discretize, libscpb), Int32,
        (Ptr{Cvoid}, Int32, Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Ptr{Int32}, Ptr{Float64}),
        h.ptr, 0, B, N, Nsub, t_grid, xd, ud, p, iSx, feas_tol, A, Bm, Bp, F, r, E, defect, feas, secs), "scpb_discretize")
    return secs[]
end

# propagate(sol, pbm; res) (src/solvers/discretization.jl:515-562, FOH): returns the nx x res matrix of xc values
function propagate(h::Handle, N, res, t_grid, xd, ud, p)
    xc = Matrix{Float64}(undef, size(xd, 1), res)
    secs = Ref{Float64}(0.0)
    check(h, ccall((:scpb_propagate, libscpb), Int32,
        (Ptr{Cvoid}, Int32, Int32, Int32, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}), h.ptr, 0, 1, N, res, t_grid, xd, ud, p, xc, secs), "scpb_propagate")
    return xc
end

struct ConeOpts
    feastol::Float64; abstol::Float64; reltol::Float64; delta::Float64; delta_dyn::Float64
    maxit::Int32; nref::Int32; verbose::Int32; group::Int32; equil::Int32; threads::Int32; lanes::Int32
end   # field order and types mirror scpb_cone_opts (include/scpb.h)
ConeOpts(; maxit = 0, verbose = 0) = ConeOpts(0, 0, 0, 0, 0, maxit, -1, verbose, 0, -1, 0, 0)

function cone_setup(h::Handle, n, p, m, A_rp, A_ci, G_rp, G_ci, l, soc_dims::Vector{Int32}, perm)
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(h, ccall((:scpb_cone_setup, libscpb), Int32,
        (Ptr{Cvoid}, Int32, Int32, Int32, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Ptr{Int32}, Int32, Int32, Ptr{Int32},
         Ptr{Int32}, Ptr{Ptr{Cvoid}}), h.ptr, n, p, m, A_rp, A_ci, G_rp, G_ci, l, length(soc_dims), soc_dims, perm, out),
        "scpb_cone_setup")
    return out[]
end

function cone_solve(h::Handle, cone, B, Av, Gv, c, b, hh, opts::ConeOpts, x, y, z, s, pobj, dobj, status, iters)
    secs = Ref{Float64}(0.0)
    check(h, ccall((:scpb_cone_solve, libscpb), Int32,
        (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{ConeOpts},
         Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Int32},
         Ptr{Float64}), cone, B, Av, Gv, c, b, hh, opts, x, y, z, s, pobj, dobj, status, iters, secs), "scpb_cone_solve")
    return secs[]
end

cone_free(cone) = ccall((:scpb_cone_free, libscpb), Int32, (Ptr{Cvoid},), cone)

# scpb_scvx_attach / scpb_scvx_solve (SCvx.solve for a batch, src/solvers/scvx.jl:460-546): the descriptor mirrors
# scpb_scvx_desc; Q holds the rows of the original cost and of g_ic / g_tc over the scaled solver variables
struct ScvxDesc
    lam::Float64; rho_0::Float64; rho_1::Float64; rho_2::Float64; beta_sh::Float64; beta_gr::Float64
    eta_init::Float64; eta_lb::Float64; eta_ub::Float64
    oeta::Int32; n_ic::Int32; n_tc::Int32; reserved::Int32
end

function scvx_attach(h::Handle, ptr, desc::ScvxDesc, Q_rp::Vector{Int32}, Q_ci::Vector{Int32}, Q_v::Vector{Float64},
                     Q_c::Vector{Float64})
    check(h, ccall((:scpb_scvx_attach, libscpb), Int32,
        (Ptr{Cvoid}, Ref{ScvxDesc}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}), ptr, desc, Q_rp, Q_ci, Q_v, Q_c),
        "scpb_scvx_attach")
end

function scvx_solve(h::Handle, ptr, B, xd0, ud0, p0, opts::ConeOpts, xd, ud, p, status, iters, J, dev, feas, eta, timing)
    check(h, ccall((:scpb_scvx_solve, libscpb), Int32,
        (Ptr{Cvoid}, Int32, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{ConeOpts}, Ptr{Float64}, Ptr{Float64},
         Ptr{Float64}, Ptr{Int32}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}, Ptr{Int32}, Ptr{Float64}, Ptr{Float64}),
        ptr, B, xd0, ud0, p0, opts, xd, ud, p, status, iters, J, dev, feas, eta, timing), "scpb_scvx_solve")
end

# scpb_ptr_setup / scpb_ptr_solve / scpb_ptr_free: same pattern; the descriptor struct mirrors scpb_ptr_desc
# field by field (27 Int32 + 3 Float64).  See INTEGRATION.md section 3.

end # module
