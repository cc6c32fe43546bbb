# SFBB200.jl — reference-side binding of libsfb_b200.so (include/sfb_b200.h).
#
# NOT EXECUTED IN THIS REPOSITORY'S CI: the build image has no Julia.  It is the thin `ccall` layer a
# maintainer of hsgg/SphericalFourierBesselDecompositions.jl would add so that `calc_Wr_lm` and
# `power_win_mix` keep their signatures (src/windows.jl:528,540,750,751,781,809,994) and run on the GPU.
# Argument marshalling is deliberately trivial: column-major arrays, Int64 sizes, ComplexF64 interleaved.
module SFBB200

using SparseArrays
using LinearAlgebra: I, UniformScaling
import SphericalFourierBesselDecompositions as SFB
using SphericalFourierBesselDecompositions: AnlmModes, ClnnModes, ClnnBinnedModes, ConfigurationSpaceModes,
    SeparableArray, getlmsize, getlnnsize, window_r
using Healpix: Alm

const libsfb = get(ENV, "SFB_B200_LIB", "libsfb_b200.so")

function check(status::Integer)
    status == 0 && return nothing
    msg = unsafe_string(ccall((:sfb_last_error, libsfb), Cstring, ()))
    error(msg)                      # ErrorException, like error()/@assert in the reference
end

# Multi-GPU: ONE call.  After `set_devices(n)` power_win_mix / calc_Wr_lm shard over GPUs 0..n-1 inside the library
# (it replaces the pmap gather of src/windows.jl:834-861); nothing else in the shim changes.
set_devices(n::Integer) = check(ccall((:sfb_set_devices, libsfb), Int32, (Int32,), n))

# Result arrays in page-locked memory (sfb_host_alloc), so that every GPU copies its column slab at full PCIe speed;
# freed by a finalizer.  A plain `Matrix{Float64}(undef, ...)` works too (pageable: on one GPU the library stages the copies
# itself through a pinned ring and a host thread pool, on several GPUs the driver does).
function pinned_matrix(::Type{T}, dims::Integer...) where {T}
    p = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:sfb_host_alloc, libsfb), Int32, (Ptr{Ptr{Cvoid}}, Int64), p, max(1, prod(dims) * sizeof(T))))
    A = unsafe_wrap(Array, Ptr{T}(p[]), dims; own=false)
    finalizer(_ -> ccall((:sfb_host_free, libsfb), Int32, (Ptr{Cvoid},), p[]), A)
    return A
end

# r .* √Δr .* precompute_gnlr(amodes, wmodes)   (src/windows.jl:799) — stays in Julia (input of the path).
# precompute_gnlr allocates `fill(NaN, nr, size(amodes.basisfunctions.knl)...)` (src/windows.jl:551); for
# AnlmModes(kmax, rmin, rmax) that knl table is the UNTRIMMED one (SphericalBesselGNLs.jl:313-316), larger than
# amodes.nmax x (amodes.lmax+1) (src/modes.jl:144,156).  The C ABI indexes G[r + nr*(n + nmax*l)] with
# nmax = amodes.nmax, so the table is trimmed to exactly nr x amodes.nmax x (amodes.lmax+1) here.
function rsdrgnlr(amodes, wmodes)
    r, Δr = window_r(wmodes)
    G = r .* .√Δr .* SFB.Windows.precompute_gnlr(amodes, wmodes)
    return Array(G[:, 1:amodes.nmax, 1:amodes.lmax+1])
end

############################## calc_Wr_lm ##############################

# src/windows.jl:528-537
function calc_Wr_lm(win::AbstractMatrix{Float64}, LMAX::Integer, Wnside::Integer; niter=3)
    win = win isa Matrix{Float64} ? win : Matrix{Float64}(win)
    nr, npix = size(win)
    Wr_lm = Matrix{ComplexF64}(undef, nr, getlmsize(LMAX))
    GC.@preserve win Wr_lm check(ccall((:sfb_calc_wr_lm, libsfb), Int32,
        (Ptr{Float64}, Int64, Int64, Int64, Int64, Int64, Int64, Int32, Ptr{ComplexF64}),
        win, nr, npix, stride(win, 2), Wnside, LMAX, niter, 0, Wr_lm))
    return Wr_lm
end

# src/windows.jl:540-545
function calc_Wr_lm(win::SeparableArray, LMAX::Integer, Wnside::Integer; niter=3)
    mask = Vector{Float64}(win.mask)
    wlm = Alm(LMAX, LMAX)
    GC.@preserve mask wlm check(ccall((:sfb_calc_wlm_mask, libsfb), Int32,
        (Ptr{Float64}, Int64, Int64, Int64, Int64, Ptr{ComplexF64}),
        mask, length(mask), Wnside, LMAX, niter, wlm.alm))
    return SeparableArray(win.phi, wlm, name1=:phi, name2=:wlm)
end

############################## win_lnn ##############################

# src/windows.jl:382-391 (with calc_intr_gg_fn :394-418): shot-noise window from Wr_00 of the same stage 1
function win_lnn(win::AbstractMatrix{Float64}, wmodes::ConfigurationSpaceModes, cmodes::ClnnModes)
    win = win isa Matrix{Float64} ? win : Matrix{Float64}(win)
    amodes = cmodes.amodes
    G = rsdrgnlr(amodes, wmodes)
    lnn = cmodes.lnn
    nr, npix = size(win)
    Wlnn = Vector{Float64}(undef, getlnnsize(cmodes))
    GC.@preserve win G lnn Wlnn check(ccall((:sfb_win_lnn, libsfb), Int32,
        (Ptr{Float64}, Int64, Int64, Int64, Int64, Ptr{Float64}, Int64, Int64, Ptr{Int64}, Int64, Ptr{Float64}),
        win, nr, npix, stride(win, 2), amodes.nside, G, amodes.nmax, amodes.lmax, lnn, size(lnn, 2), Wlnn))
    return Wlnn
end

############################## power_win_mix ##############################

power_win_mix(win, wmodes::ConfigurationSpaceModes, cmodes::ClnnModes; kwargs...) =
    power_win_mix(win, win, wmodes, cmodes; kwargs...)                                   # src/windows.jl:750
power_win_mix(win, w̃, v, wmodes::ConfigurationSpaceModes, bcmodes::ClnnBinnedModes; kwargs...) =
    power_win_mix(win, win, w̃, v, wmodes, bcmodes; kwargs...)                            # src/windows.jl:751

# src/windows.jl:781-805
function power_win_mix(win1::AbstractMatrix{Float64}, win2::AbstractMatrix{Float64},
                       wmodes::ConfigurationSpaceModes, cmodes::ClnnModes;
                       div2Lp1=false, interchange_NN′=false, lnn_min=1)
    amodes = cmodes.amodes
    G = rsdrgnlr(amodes, wmodes)
    lnn = cmodes.lnn
    lnnsize = getlnnsize(cmodes)
    n = lnnsize - lnn_min + 1
    mix = pinned_matrix(Float64, n, n)
    w1 = win1 isa Matrix{Float64} ? win1 : Matrix{Float64}(win1)
    w2 = win2 === win1 ? w1 : (win2 isa Matrix{Float64} ? win2 : Matrix{Float64}(win2))
    nr, npix = size(w1)
    p2 = win2 === win1 ? Ptr{Float64}(C_NULL) : pointer(w2)
    GC.@preserve w1 w2 G lnn mix check(ccall((:sfb_power_win_mix, libsfb), Int32,
        (Ptr{Float64}, Ptr{Float64}, Int64, Int64, Int64, Int64, Ptr{Float64}, Int64, Int64, Ptr{Int64}, Int64,
         Int64, Int32, Int32, Ptr{Float64}),
        w1, p2, nr, npix, stride(w1, 2), amodes.nside, G, amodes.nmax, amodes.lmax, lnn, lnnsize, lnn_min,
        div2Lp1, interchange_NN′, mix))
    return mix
end

# src/windows.jl:809-814
function power_win_mix(win1::SeparableArray, win2::SeparableArray, wmodes::ConfigurationSpaceModes,
                       cmodes::ClnnModes; kwargs...)
    bcmodes = ClnnBinnedModes(I, I, cmodes)
    return power_win_mix(win1, win2, I, I, wmodes, bcmodes; kwargs...)
end

# SparseMatrixCSC -> (colptr, rowval, nzval, size); UniformScaling -> NULLs (src/windows.jl:829-830)
csc(m::UniformScaling, n, dim) = (Ptr{Int64}(C_NULL), Ptr{Int64}(C_NULL), Ptr{Float64}(C_NULL), n, nothing)
function csc(m::AbstractMatrix, n, dim)
    s = SparseMatrixCSC{Float64,Int64}(sparse(m))
    return (pointer(s.colptr), pointer(s.rowval), pointer(s.nzval), size(s, dim), s)
end

# src/windows.jl:994-1015 (dense, :825-862) and the separable specialisation (:942-990)
function power_win_mix(win1, win2, w̃mat, vmat, wmodes::ConfigurationSpaceModes, bcmodes::ClnnBinnedModes;
                       div2Lp1=false, interchange_NN′=false)
    cmodes = bcmodes.cmodes
    amodes = cmodes.amodes
    G = rsdrgnlr(amodes, wmodes)
    lnn = cmodes.lnn
    lnnsize = getlnnsize(cmodes)
    wc, wr, wv, LNN1, wkeep = csc(w̃mat, lnnsize, 1)
    vc, vr, vv, LNN2, vkeep = csc(vmat, lnnsize, 2)
    mix = Matrix{Float64}(undef, LNN1, LNN2)
    if win1 isa SeparableArray
        phi = Vector{Float64}(win1.phi); mask = Vector{Float64}(win1.mask)
        GC.@preserve phi mask G lnn wkeep vkeep mix check(ccall((:sfb_power_win_mix_separable, libsfb), Int32,
            (Ptr{Float64}, Ptr{Float64}, Int64, Int64, Int64, Ptr{Float64}, Int64, Int64, Ptr{Int64}, Int64,
             Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int64, Int32, Int32,
             Ptr{Float64}),
            phi, mask, length(phi), length(mask), amodes.nside, G, amodes.nmax, amodes.lmax, lnn, lnnsize,
            wc, wr, wv, LNN1, vc, vr, vv, LNN2, div2Lp1, interchange_NN′, mix))
    else
        w1 = win1 isa Matrix{Float64} ? win1 : Matrix{Float64}(win1)
        nr, npix = size(w1)
        # like the reference, the second transform is also taken from win1 (src/windows.jl:1005-1006)
        GC.@preserve w1 G lnn wkeep vkeep mix check(ccall((:sfb_power_win_mix_binned, libsfb), Int32,
            (Ptr{Float64}, Int64, Int64, Int64, Int64, Ptr{Float64}, Int64, Int64, Ptr{Int64}, Int64,
             Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int64, Int32, Int32,
             Ptr{Float64}),
            w1, nr, npix, stride(w1, 2), amodes.nside, G, amodes.nmax, amodes.lmax, lnn, lnnsize,
            wc, wr, wv, LNN1, vc, vr, vv, LNN2, div2Lp1, interchange_NN′, mix))
    end
    return mix
end

############################## §8f rows: calc_wmix, SFB transforms, deconvolution ##############################

# src/windows.jl:299-364
function calc_wmix(win::AbstractMatrix{Float64}, wmodes::ConfigurationSpaceModes, amodes::AnlmModes; neg_m=false)
    win = win isa Matrix{Float64} ? win : Matrix{Float64}(win)
    G = rsdrgnlr(amodes, wmodes)
    nr, npix = size(win)
    nlmsize = SFB.getnlmsize(amodes)
    wmix = pinned_matrix(ComplexF64, nlmsize, nlmsize)
    nmax_l = Vector{Int64}(amodes.nmax_l); lmax_n = Vector{Int64}(amodes.lmax_n)
    GC.@preserve win G nmax_l lmax_n wmix check(ccall((:sfb_calc_wmix, libsfb), Int32,
        (Ptr{Float64}, Int64, Int64, Int64, Int64, Ptr{Float64}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Int32, Ptr{ComplexF64}),
        win, nr, npix, stride(win, 2), amodes.nside, G, amodes.nmax, amodes.lmax, nmax_l, lmax_n, neg_m, wmix))
    return wmix
end
calc_wmix_all(win::AbstractMatrix{Float64}, wmodes, amodes) =                      # src/window_chains.jl:578-582
    (calc_wmix(win, wmodes, amodes), calc_wmix(win, wmodes, amodes, neg_m=true))

# radial tables trimmed like rsdrgnlr: g_nl(r) and g_nl(r) r² Δr
function gnlr_tables(amodes, wmodes)
    r, Δr = window_r(wmodes)
    g = Array(SFB.Windows.precompute_gnlr(amodes, wmodes)[:, 1:amodes.nmax, 1:amodes.lmax+1])
    return g, g .* (r .^ 2 .* Δr)
end

# src/cat2anlm.jl:326-373
function field2anlm(f_xyz::AbstractMatrix{Float64}, wmodes::ConfigurationSpaceModes, amodes::AnlmModes)
    f = f_xyz isa Matrix{Float64} ? f_xyz : Matrix{Float64}(f_xyz)
    _, T = gnlr_tables(amodes, wmodes)
    out = Vector{ComplexF64}(undef, SFB.getnlmsize(amodes))
    nmax_l = Vector{Int64}(amodes.nmax_l); lmax_n = Vector{Int64}(amodes.lmax_n)
    GC.@preserve f T nmax_l lmax_n out check(ccall((:sfb_field2anlm, libsfb), Int32,
        (Ptr{Float64}, Int64, Int64, Int64, Ptr{Float64}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{ComplexF64}),
        f, size(f, 1), size(f, 2), stride(f, 2), T, amodes.nmax, amodes.lmax, nmax_l, lmax_n, out))
    return out
end

# src/cat2anlm.jl:385-422
function anlm2field(f_nlm::AbstractVector{ComplexF64}, wmodes::ConfigurationSpaceModes, amodes::AnlmModes)
    g, _ = gnlr_tables(amodes, wmodes)
    f = Vector{ComplexF64}(f_nlm)
    out = pinned_matrix(Float64, wmodes.nr, wmodes.npix)
    nmax_l = Vector{Int64}(amodes.nmax_l); lmax_n = Vector{Int64}(amodes.lmax_n)
    GC.@preserve f g nmax_l lmax_n out check(ccall((:sfb_anlm2field, libsfb), Int32,
        (Ptr{ComplexF64}, Ptr{Float64}, Int64, Int64, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int64),
        f, g, wmodes.nr, amodes.nside, amodes.nmax, amodes.lmax, nmax_l, lmax_n, out, wmodes.nr))
    return out
end

# src/windows.jl:244-256 (the SeparableArray method :259-270 stays in Julia)
function win_rhat_ln(win::AbstractMatrix{Float64}, wmodes::ConfigurationSpaceModes, amodes::AnlmModes)
    w = win isa Matrix{Float64} ? win : Matrix{Float64}(win)
    _, T = gnlr_tables(amodes, wmodes)
    out = pinned_matrix(Float64, size(w, 2) * (amodes.lmax + 1) * amodes.nmax)
    nmax_l = Vector{Int64}(amodes.nmax_l); lmax_n = Vector{Int64}(amodes.lmax_n)
    GC.@preserve w T nmax_l lmax_n out check(ccall((:sfb_win_rhat_ln, libsfb), Int32,
        (Ptr{Float64}, Int64, Int64, Int64, Ptr{Float64}, Int64, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}),
        w, size(w, 1), size(w, 2), stride(w, 2), T, amodes.nmax, amodes.lmax, nmax_l, lmax_n, out))
    return reshape(out, size(w, 2), amodes.lmax + 1, amodes.nmax)
end

# C = bcmix \ (w̃mat * Cobs) without bringing bcmix to the host  (docs/src/tutorial_catalog.md:93-97)
function power_win_mix_solve(win::AbstractMatrix{Float64}, w̃mat, vmat, wmodes::ConfigurationSpaceModes,
                             bcmodes::ClnnBinnedModes, B::AbstractVecOrMat{Float64}; div2Lp1=false, interchange_NN′=false)
    cmodes = bcmodes.cmodes
    amodes = cmodes.amodes
    G = rsdrgnlr(amodes, wmodes)
    lnn = cmodes.lnn
    lnnsize = getlnnsize(cmodes)
    wc, wr, wv, LNN1, wkeep = csc(w̃mat, lnnsize, 1)
    vc, vr, vv, LNN2, vkeep = csc(vmat, lnnsize, 2)
    w1 = win isa Matrix{Float64} ? win : Matrix{Float64}(win)
    Bm = Matrix{Float64}(reshape(B, LNN1, :))
    X = similar(Bm)
    nr, npix = size(w1)
    GC.@preserve w1 G lnn wkeep vkeep Bm X check(ccall((:sfb_power_win_mix_binned_solve, libsfb), Int32,
        (Ptr{Float64}, Int64, Int64, Int64, Int64, Ptr{Float64}, Int64, Int64, Ptr{Int64}, Int64,
         Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int64, Ptr{Int64}, Ptr{Int64}, Ptr{Float64}, Int64, Int32, Int32,
         Ptr{Float64}, Int64, Ptr{Float64}, Ptr{Float64}),
        w1, nr, npix, stride(w1, 2), amodes.nside, G, amodes.nmax, amodes.lmax, lnn, lnnsize,
        wc, wr, wv, LNN1, vc, vr, vv, LNN2, div2Lp1, interchange_NN′, Bm, size(Bm, 2), X, C_NULL))
    return B isa AbstractVector ? vec(X) : X
end

end # module
