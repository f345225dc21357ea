# SeveroB200.jl — Julia overlay that routes Severo.jl's IRLBA-PCA hot path through libsevero_b200.so.
# Same function names, keyword arguments and return types as the reference (src/normalize.jl:40-79,
# src/scaling.jl:335-357, src/irlba.jl:47-99, src/embedding.jl:46-94); only the bodies change: every loop /
# the `ccall(("irlba", libcell), ...)` becomes a `ccall` into the C ABI of include/severo_b200.h.
# NOTE: Julia is not installed in the build image, so this file is the binding a maintainer would add; it is
# exercised through the identical C ABI by the Python ctypes harness (severo.jl_b200/api.py).
module SeveroB200

using SparseArrays, LinearAlgebra, NamedArrays, Random, Statistics
import Severo
import Severo: CenteredMatrix, NamedCenteredMatrix, NamedCountMatrix, LinearEmbedding  # Severo re-exports Distances' Euclidean / CosineDist

const libsvb = get(ENV, "SEVERO_B200_LIB", joinpath(@__DIR__, "..", "severo.jl_b200", "libsevero_b200.so"))
const SVB_I32, SVB_I64, SVB_F32, SVB_F64 = Cint(0), Cint(1), Cint(2), Cint(3)
svbtype(::Type{Int32}) = SVB_I32; svbtype(::Type{Int64}) = SVB_I64
svbtype(::Type{Float32}) = SVB_F32; svbtype(::Type{Float64}) = SVB_F64

function check(rc::Cint)
    rc == 0 && return nothing
    msg = unsafe_string(ccall((:svb_last_error, libsvb), Cstring, ()))
    error("severo_b200 [$rc]: $msg")
end

__init__() = check(ccall((:svb_init, libsvb), Cint, (Cint,), parse(Cint, get(ENV, "LOCAL_RANK", "0"))))

# ---- device handles -------------------------------------------------------------------------------
mutable struct DeviceMatrix
    h::Ptr{Cvoid}
    function DeviceMatrix(h)
        x = new(h)
        finalizer(d -> ccall((:svb_matrix_free, libsvb), Cint, (Ptr{Cvoid},), d.h), x)
    end
end

# GC safety: handles are passed to `ccall` as the OBJECT, never as the raw field `d.h`. `ccall` keeps every argument it
# converts (Base.cconvert) rooted until the foreign call returns, so the finalizer cannot free the device matrix / operator
# while a kernel that reads it is still running; passing `d.h` would leave `d` unreferenced after the field load.
Base.unsafe_convert(::Type{Ptr{Cvoid}}, d::DeviceMatrix) = d.h

# SparseMatrixCSC{T,Int64}: colptr / rowval are 1-based Int64 — passed as they are (index_base = 1)
function upload(A::SparseMatrixCSC{T,Int64}) where {T<:Union{Int32,Int64,Float32,Float64}}
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:svb_csc_upload, libsvb), Cint,
        (Int64, Int64, Ptr{Int64}, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cint, Cint, Ref{Ptr{Cvoid}}),
        size(A, 1), size(A, 2), A.colptr, A.rowval, SVB_I64, A.nzval, svbtype(T), 1, h))
    DeviceMatrix(h[])
end

function download_values(d::DeviceMatrix, ::Type{R}, nnz::Integer) where {R}
    nz = Vector{R}(undef, nnz)
    check(ccall((:svb_matrix_download, libsvb), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Cvoid}, Cint, Cint),
        d, C_NULL, C_NULL, nz, svbtype(R), 1))
    nz
end

# the whole matrix back as a SparseMatrixCSC{Int64,Int64} (counts) — 1-based indices written by the library
function download(d::DeviceMatrix)
    nrow = Ref{Int64}(0); ncol = Ref{Int64}(0); nz = Ref{Int64}(0); vt = Ref{Cint}(0)
    check(ccall((:svb_matrix_info, libsvb), Cint, (Ptr{Cvoid}, Ref{Int64}, Ref{Int64}, Ref{Int64}, Ref{Cint}), d, nrow, ncol, nz, vt))
    colptr = Vector{Int64}(undef, ncol[] + 1); rowval = Vector{Int64}(undef, nz[]); nzval = Vector{Int64}(undef, nz[])
    check(ccall((:svb_matrix_download, libsvb), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Cvoid}, Cint, Cint),
        d, colptr, rowval, nzval, SVB_I64, 1))
    SparseMatrixCSC(nrow[], ncol[], colptr, rowval, nzval)
end

# ---- normalize.jl:40-79 ------------------------------------------------------------------------------
function normalize_cells(X::SparseMatrixCSC{<:Integer,Int64}; method=:lognormalize, scale_factor::Real=1., dtype::Type{T}=Float64) where {T<:AbstractFloat}
    method = Symbol(method)
    code = method == :lognormalize ? Cint(0) : method == :relativecounts ? Cint(1) : error("unknown normalization method: $method")
    d = upload(X)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:svb_normalize, libsvb), Cint, (Ptr{Cvoid}, Cint, Float64, Cint, Ref{Ptr{Cvoid}}),
        d, code, Float64(convert(dtype, scale_factor)), svbtype(dtype), h))
    out = DeviceMatrix(h[])
    SparseMatrixCSC(size(X, 1), size(X, 2), copy(X.colptr), copy(X.rowval), download_values(out, dtype, nnz(X)))
end

function normalize_cells(X::NamedCountMatrix; kw...)
    S = normalize_cells(X.array; kw...)
    NamedArray(S, X.dicts, X.dimnames)                       # normalize.jl:77-78
end

# ---- scaling.jl:119-147, 199-217, 335-357 ------------------------------------------------------------
function mean_var(A::SparseMatrixCSC)
    d = upload(A); mu = zeros(size(A, 2)); var = zeros(size(A, 2))
    check(ccall((:svb_mean_var, libsvb), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), d, mu, var))
    mu, var
end

function scale_data(A::SparseMatrixCSC, scale_max::R=Inf) where {R<:AbstractFloat}
    d = upload(A); mu = zeros(Float64, size(A, 2)); h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:svb_scale, libsvb), Cint, (Ptr{Cvoid}, Float64, Cint, Ref{Ptr{Cvoid}}, Ptr{Float64}),
        d, Float64(scale_max), svbtype(R), h, mu))
    out = DeviceMatrix(h[])
    B = SparseMatrixCSC(size(A, 1), size(A, 2), copy(A.colptr), copy(A.rowval), download_values(out, R, nnz(A)))
    B, convert(Vector{R}, mu)
end

function scale_features(X::NamedArray{T,2,SparseMatrixCSC{T,Int64}}; scale_max::Real=Inf,
        dtype::Type{<:AbstractFloat}=(T <: AbstractFloat ? T : Float64), features=nothing) where {T}
    features !== nothing && (X = X[:, features])
    B, mu = scale_data(X.array, convert(dtype, scale_max))
    CenteredMatrix(NamedArray(B, X.dicts, X.dimnames), NamedArray(mu, (X.dicts[2],), (X.dimnames[2],)))   # scaling.jl:344
end

# ---- variablefeatures.jl:19-50,128-161 (:vst): the two data sweeps on the device, loess + top-k on the host ----------
import Loess: loess, predict

function standardized_var_clipped(A::SparseMatrixCSC{<:Integer}, mu::Vector{Float64}, sd::Vector{Float64}; vmax=sqrt(size(A, 1)))
    d = upload(A); out = zeros(size(A, 2))
    check(ccall((:svb_stdvar_clipped, libsvb), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Float64, Ptr{Float64}),
        d, mu, sd, Float64(vmax), out))
    out
end

# filtering.jl:15-35,101-106 — cells first, then features on the remaining cells; returns (A[CI, FI], CI, FI)
function filter_counts(A::SparseMatrixCSC{<:Integer}; min_cells=0, min_features=0, min_feature_count=0, min_umi=0)
    d = upload(A)
    CI, FI = zeros(UInt8, size(A, 1)), zeros(UInt8, size(A, 2))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:svb_filter_counts, libsvb), Cint,
        (Ptr{Cvoid}, Int64, Int64, Int64, Int64, Ptr{UInt8}, Ptr{UInt8}, Ref{Ptr{Cvoid}}),
        d, min_cells, min_features, min_feature_count, min_umi, CI, FI, h))
    download(DeviceMatrix(h[])), CI .!= 0, FI .!= 0
end
function filter_counts(A::NamedCountMatrix; kw...)
    counts, CI, FI = filter_counts(A.array; kw...)
    barcodes, features = names(A)
    NamedArray(counts, (barcodes[CI], features[FI]), A.dimnames)
end

function variance_stabilizing_transformation(A::SparseMatrixCSC{<:Integer}; loess_span::Real=0.5)
    mu, var = mean_var(A)                                        # device, order-exact Welford (scaling.jl:18-34)
    sd = sqrt.(var)
    non_const = sd .> 0
    xs, ys = log10.(mu[non_const]), log10.(sd[non_const])
    model = loess(xs, ys, span=loess_span)                       # host, O(genes) (variablefeatures.jl:41)
    expected = copy(sd)
    expected[non_const] = 10 .^ predict(model, xs)
    expected .= ifelse.(isnan.(expected), 0.0, expected)         # nan2zero! (variablefeatures.jl:30-32)
    standardized_var_clipped(A, mu, expected)                    # device (variablefeatures.jl:21-28)
end

# the selectors next to :vst (variablefeatures.jl:52-103,135-155): row_norm(counts, 1) and its per-gene moments on the
# device (svb_normalize / svb_mean_var, svb_row_sums for :saunders), the gene-length arithmetic by Severo's own functions
function _log_VMR(norm::SparseMatrixCSC)
    mu, var = mean_var(norm)                                     # device; same Welford as mean_var(T, identity, A) scaling.jl:157-187
    log1p.(mu), log.(var ./ mu)                                  # scaling.jl:190-193
end

function _variable_feature_metric(counts::SparseMatrixCSC{<:Integer}, method::Symbol; kw...)
    norm = if :norm in keys(kw)
        n = kw[:norm]; isa(n, NamedArray) ? n.array : n
    else
        normalize_cells(counts; method=:relativecounts, scale_factor=1.0)        # row_norm(counts.array, one(dtype)) :140
    end
    if method == :saunders
        ncells, ngenes = size(counts)
        trx = zeros(Int64, ncells)
        d = upload(counts)
        check(ccall((:svb_row_sums, libsvb), Cint, (Ptr{Cvoid}, Ptr{Int64}), d, trx))
        mu, var = mean_var(norm)
        nolan = Statistics.mean(1 ./ trx)
        alpha = get(kw, :alpha_thresh, 0.1) / ngenes
        upper = mu .+ Severo.qnorm(1 - alpha / 2) .* sqrt.(mu .* nolan ./ ncells)
        metric = zeros(ngenes)
        J = (var ./ nolan) .> upper
        metric[J] = log10.(var[J]) .- log10.(mu[J] .* nolan)
        metric
    elseif method == :dispersion
        Severo.nan2zero!(_log_VMR(norm)[2])
    elseif method == :meanvarplot
        mu, disp = _log_VMR(norm)
        mu = Severo.nan2zero!(mu); disp = Severo.nan2zero!(disp)
        num_bins = get(kw, :num_bins, 20)
        _, bins = Severo.cut(mu, num_bins; method=get(kw, :binning_method, :width))
        bin_mean, bin_std = Severo.mean_std(disp, bins, num_bins)
        Severo.nan2zero!((disp .- bin_mean[bins]) ./ bin_std[bins])
    else
        error("unknown selection method: $method")
    end
end

function find_variable_features(counts::NamedCountMatrix, nfeatures=2000; method=:vst, kw...)
    method = Symbol(method)
    metric = method == :vst ? variance_stabilizing_transformation(counts.array; kw...) :
                              _variable_feature_metric(counts.array, method; kw...)
    selected = partialsortperm(metric, 1:nfeatures, rev=true)    # variablefeatures.jl:159
    NamedArray(selected, (names(counts, 2)[selected],), (dimnames(counts, 2),))
end

# ---- the operator + irlba.jl:47-99 -----------------------------------------------------------------------
mutable struct DeviceOperator
    h::Ptr{Cvoid}
    function DeviceOperator(h)
        x = new(h)
        finalizer(o -> ccall((:svb_operator_free, libsvb), Cint, (Ptr{Cvoid},), o.h), x)
    end
end

Base.unsafe_convert(::Type{Ptr{Cvoid}}, o::DeviceOperator) = o.h   # rooted by ccall for the whole call (see DeviceMatrix)

_mu_ptr(mu::Nothing) = Ptr{Float64}(C_NULL)
_mu_ptr(mu::AbstractVector) = convert(Vector{Float64}, mu)

function operator(A::SparseMatrixCSC, mu=nothing; transposed::Bool=false)
    d = upload(A isa SparseMatrixCSC{<:AbstractFloat} ? A : convert(SparseMatrixCSC{Float64,Int64}, A))
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:svb_operator_create, libsvb), Cint, (Ptr{Cvoid}, Ptr{Float64}, Cint, Ref{Ptr{Cvoid}}),
        d, _mu_ptr(mu), Cint(transposed), h))
    DeviceOperator(h[])
end
operator(A::Adjoint{<:Any,<:SparseMatrixCSC}, mu=nothing) = operator(parent(A), mu; transposed=true)   # test_irlba.jl:111
function operator(A::StridedMatrix{Float64}, mu=nothing)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:svb_operator_create_dense, libsvb), Cint,
        (Int64, Int64, Ptr{Float64}, Int64, Ptr{Float64}, Cint, Ref{Ptr{Cvoid}}),
        size(A, 1), size(A, 2), A, stride(A, 2), _mu_ptr(mu), Cint(0), h))
    DeviceOperator(h[])
end
# The same operator built from the RAW COUNTS of the HVG columns (svb_operator_create_counts): fuses
# normalize_cells(:lognormalize) -> Y[:, hvf] -> scale_features(scale_max) (normalize.jl:40-55, scaling.jl:335-357);
# the scaled matrix is never materialised. `libsize` = vec(sum(counts_all_genes, dims=2)) (normalize.jl:24).
# moments = :exact uses the reference's sequential Welford (mean_var of the normalised columns, scaling.jl:18-34),
# :fast lets the library compute them in two parallel passes (a few ulp away; spans all ranks of a communicator).
function counts_operator(counts_hvg::SparseMatrixCSC{<:Integer}, libsize::Vector{Int64};
                         scale_factor::Real=1e4, scale_max::Real=Inf, moments::Symbol=:exact, levels::Integer=0)
    n = size(counts_hvg, 2)
    mean = var = Ptr{Float64}(C_NULL)
    if moments == :exact && counts_hvg isa SparseMatrixCSC{Int32,Int64}
        # the pipelined entry point: the order-exact moments are computed while the matrix crosses PCIe (bit-identical)
        d, mean, var = upload_lognorm_moments(counts_hvg, libsize; scale_factor)
    elseif moments == :exact
        d = upload(counts_hvg)
        y = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:svb_normalize_libsize, libsvb), Cint, (Ptr{Cvoid}, Ptr{Int64}, Cint, Float64, Cint, Ref{Ptr{Cvoid}}),
            d, libsize, 0, scale_factor, 3, y))
        mean, var = zeros(n), zeros(n)
        check(ccall((:svb_mean_var, libsvb), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}), y[], mean, var))
        ccall((:svb_matrix_free, libsvb), Cint, (Ptr{Cvoid},), y[])
    else
        d = upload(counts_hvg)
    end
    mu = zeros(n)                                    # receives mean/std, the stored centre (scaling.jl:207)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:svb_operator_create_counts, libsvb), Cint,
        (Ptr{Cvoid}, Ptr{Int64}, Float64, Ptr{Float64}, Ptr{Float64}, Float64, Cint, Ptr{Float64}, Ref{Ptr{Cvoid}}),
        d, libsize, scale_factor, mean, var, scale_max, levels, mu, h))
    DeviceOperator(h[]), mu
end

# upload of the HVG counts + the order-exact moments of their log-normalised values, pipelined (densest genes first; the Welford
# chains run on side streams during the upload): (DeviceMatrix, mean, var), bit-identical to normalize -> mean_var
function upload_lognorm_moments(counts_hvg::SparseMatrixCSC{Int32,Int64}, libsize::Vector{Int64}; scale_factor::Real=1e4)
    n = size(counts_hvg, 2)
    mean = zeros(n); var = zeros(n); h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:svb_csc_upload_lognorm_moments, libsvb), Cint,
        (Int64, Int64, Ptr{Int64}, Ptr{Cvoid}, Cint, Ptr{Int32}, Cint, Ptr{Int64}, Float64, Ptr{Float64}, Ptr{Float64}, Ref{Ptr{Cvoid}}),
        size(counts_hvg, 1), n, counts_hvg.colptr, counts_hvg.rowval, SVB_I64, counts_hvg.nzval, 1, libsize, scale_factor, mean, var, h))
    DeviceMatrix(h[]), mean, var
end

# embedding(counts; ...) in one call: pca of the scaled HVG matrix straight from the counts
function pca_counts(counts::SparseMatrixCSC{<:Integer}, hvf::AbstractVector{<:Integer}, npcs::Integer;
                    scale_factor::Real=1e4, scale_max::Real=Inf, moments::Symbol=:exact, tol=1e-5, init=nothing)
    libsize = convert(Vector{Int64}, vec(sum(counts, dims=2)))
    sub = counts[:, hvf]
    op, mu = counts_operator(sub, libsize; scale_factor, scale_max, moments)
    m, n = size(sub)
    U, s, V = zeros(m, npcs), zeros(npcs), zeros(n, npcs)
    init === nothing && (init = randn(n))
    iter = Ref{Int64}(0); mprod = Ref{Int64}(0)
    info = ccall((:svb_irlba, libsvb), Cint,
        (Ptr{Cvoid}, Int64, Int64, Int64, Int64, Float64, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Int64}, Ref{Int64}),
        op, npcs, min(npcs + 7, min(m, n)), 1000, 0, tol, tol, init, s, U, V, iter, mprod)
    info == 0 || error("convergence failed")
    U * Diagonal(s), s ./ sqrt(max(1, m - 1)), V, mu     # coordinates, stdev, loadings (embedding.jl:67-68), centre
end

operator(S::CenteredMatrix) = operator(Severo._A_mu(S)...)
operator(S::NamedCenteredMatrix) = operator(S.A.array, S.mu.array)

# Same signature, buffers, defaults and error as Severo.irlba! (irlba.jl:47-76); `rng` only seeds `init`
# (the breakdown vector comes from the device generator: no callback into Julia, see T6).
function irlba!(rng, A, U::Matrix{Float64}, s::Vector{Float64}, V::Matrix{Float64}; init=nothing, tol=1e-5, svtol=tol, maxit=1000, restart=0)
    m, n = size(A)
    nu = length(s)
    m_b = min(nu + 7, min(m, n))                               # irlba.jl:50-58
    init === nothing && (init = randn(rng, n))                  # irlba.jl:62-64
    op = operator(A)
    iter = Ref{Int64}(0); mprod = Ref{Int64}(0)
    info = ccall((:svb_irlba, libsvb), Cint,
        (Ptr{Cvoid}, Int64, Int64, Int64, Int64, Float64, Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Int64}, Ref{Int64}),
        op, nu, m_b, maxit, restart, tol, svtol, init, s, U, V, iter, mprod)
    info == 0 || error("convergence failed")                  # irlba.jl:73
    SVD(U, s, V')                                              # irlba.jl:75
end

function irlba(A, nu::Integer; kw...)
    m, n = size(A)
    irlba!(Random.default_rng(), A, zeros(m, nu), zeros(nu), zeros(n, nu); kw...)
end

# scaling.jl:274-296  C'C  (mul!(C, S', S, 1, 0)): the n x n Gram matrix of the centred operator, built on the device
function gram(A)
    n = size(A, 2)
    G = Matrix{Float64}(undef, n, n)
    op = operator(A)
    check(ccall((:svb_gram, libsvb), Cint, (Ptr{Cvoid}, Ptr{Float64}), op, G))
    G
end

# embedding.jl:30-44 tssvd: same keywords and return type; C'C and its eigenpairs are computed on the device
function tssvd(A::AbstractMatrix; nsv::Int=6, ritzvec::Bool=true, tol::Float64=0.0, maxiter::Int=1000, ncv::Int=2*nsv, v0=nothing)
    m, n = size(A)
    v0 === nothing && (v0 = randn(n))
    op = operator(A)
    r = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:svb_tssvd, libsvb), Cint, (Ptr{Cvoid}, Int64, Int64, Int64, Float64, Ptr{Float64}, Ref{Ptr{Cvoid}}),
        op, nsv, min(ncv, n), maxiter, tol, convert(Vector{Float64}, v0), r))
    info = Ref{Cint}(0)
    ccall((:svb_result_info, libsvb), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ptr{Int64}, Ref{Cint}),
        r[], C_NULL, C_NULL, C_NULL, C_NULL, C_NULL, info)
    Sigma = Vector{Float64}(undef, nsv); phi = Matrix{Float64}(undef, n, nsv)
    U = Matrix{Float64}(undef, m, ritzvec ? nsv : 0)                     # embedding.jl:37-42
    rc = ccall((:svb_result_download, libsvb), Cint, (Ptr{Cvoid}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Cint),
        r[], Sigma, ritzvec ? U : Ptr{Float64}(C_NULL), phi, 0)   # arrays are rooted by ccall; no bare pointer(U)
    ccall((:svb_result_free, libsvb), Cint, (Ptr{Cvoid},), r[])
    check(rc)
    info[] == 0 || error("convergence failed")
    SVD(U, Sigma, phi')                                                  # embedding.jl:44
end

# embedding.jl:46-76: algorithm = :irlba (this path), :arpack (the reference default: served by the same device solver at
# Arpack-like accuracy, maxiter/ncv mapped to maxit/work), :tssvd (Gram matrix + eigenpairs on the device)
function _pca(X, npcs::Int64; algorithm=:arpack, kw...)
    m, n = size(X)
    npcs = min(min(m, n), npcs)
    S = if algorithm == :irlba
        irlba(X, npcs; kw...)
    elseif algorithm == :arpack
        kwd = Dict{Symbol,Any}(kw)
        haskey(kwd, :maxiter) && (kwd[:maxit] = pop!(kwd, :maxiter))
        pop!(kwd, :ncv, nothing); pop!(kwd, :nsv, nothing)
        get!(kwd, :tol, 1e-10)
        irlba(X, npcs; kwd...)
    elseif algorithm == :tssvd
        tssvd(X; nsv=npcs, kw...)
    else
        error("algorithm $algorithm is outside the B200 hot path (use :irlba, :arpack or :tssvd)")
    end
    Z = view(S.U, :, 1:npcs) * Diagonal(view(S.S, 1:npcs))
    stdev = view(S.S, 1:npcs) ./ sqrt(max(1, m - 1))
    Z, stdev, S.V
end

_pca(X::NamedCenteredMatrix, npcs::Int64; kw...) = _pca(CenteredMatrix(X.A.array, X.mu.array), npcs; kw...)

# embedding.jl:81-94, 202-212 — same labels, same (sic) dimnames of `basis`
function pca(X::NamedCenteredMatrix, npcs::Int64; kw...)
    Z, stdev, loadings = _pca(X, npcs; kw...)
    k = length(stdev)
    latentnames = map(x -> string("PC-", x), 1:k)
    rownames, colnames = names(X)
    rowdim, coldim = dimnames(X)
    coordinates = NamedArray(Z, (rownames, latentnames), (rowdim, :latent))
    stdev = NamedArray(stdev, (latentnames,), (:latent,))
    basis = NamedArray(loadings, (colnames, latentnames), (rowdim, :latent))
    LinearEmbedding(X, coordinates, stdev, basis)
end

function embedding(X, ncomponents::Int64=50; method=:pca, kw...)
    Symbol(method) == :pca || error("unknown reduction method: $method")
    pca(X, ncomponents; kw...)
end

# ---- neighbours.jl:19-86 : the step after the path ------------------------------------------------------------
# Same buffers as Severo.ann! (nn_index n x k Int32, distances n x k of eltype(X), X with unit row stride); the four
# `ccall(("FindNeighbours...", libcell), ...)` become one call of the exact device search. `rng` / `ntables` drove the
# reference's randomised approximation and are accepted for signature compatibility only.
_metric(::Severo.Euclidean) = Cint(0)
_metric(::Severo.CosineDist) = Cint(1)

function ann!(rng, nn_index::Matrix{Int32}, distances::Matrix{T}, metric, X::StridedMatrix{T}, k::Int64,
        include_self::Bool, ntables::Int64) where {T<:Union{Float32,Float64}}
    n, d = size(X)
    @assert stride(X, 1) == 1
    check(ccall((:svb_knn, libsvb), Cint,
        (Ptr{Cvoid}, Cint, Int64, Int64, Int64, Int64, Cint, Cint, Cint, Ptr{Int32}, Ptr{Cvoid}),
        X, svbtype(T), n, d, stride(X, 2), k, _metric(metric), Cint(include_self), Cint(1), nn_index, distances))
    nn_index, distances
end

function ann(rng, X::AbstractMatrix{T}, k::Int64, metric=Severo.Euclidean(), include_self::Bool=true, ntables::Int64=2*size(X,2)) where {T<:AbstractFloat}
    n, d = size(X)
    ann!(rng, Matrix{Int32}(undef, n, k), Matrix{T}(undef, n, k), metric, X, k, include_self, ntables)
end

function nearest_neighbours(rng, X::NamedArray{T,2}, k::Int64; dims=:, metric=Severo.Euclidean(), include_self::Bool=true,
        ntables::Int64=2*size(X,2)) where {T}
    Z = dims === Colon() ? X.array : X.array[:, dims]
    nn_index, _ = ann(rng, Z, k, metric, include_self, ntables)
    nn = sparse(vec(nn_index'), repeat(1:size(nn_index, 1), inner=k), trues(length(nn_index)))   # neighbours.jl:79
    NamedArray(nn, (X.dicts[1], X.dicts[1]), (X.dimnames[1], X.dimnames[1]))
end
nearest_neighbours(X::NamedArray, k::Int64; kw...) = nearest_neighbours(Random.default_rng(), X, k; kw...)
nearest_neighbours(rng, em::LinearEmbedding, k::Int64; kw...) = nearest_neighbours(rng, em.coordinates, k; kw...)
nearest_neighbours(em::LinearEmbedding, k::Int64; kw...) = nearest_neighbours(Random.default_rng(), em, k; kw...)


# ---- neighbours.jl:88-152,263-270 : Jaccard index / shared nearest neighbours, the consumer of the kNN graph ------------
# `_jaccard_index` forms nn' * nn with the generic sparse product; the device computes the same entries from the neighbour
# lists directly (csrc/snn.cu) and returns them in SparseMatrixCSC order. k <= 0 selects the form without k (:96-110).
function _jaccard_index(::Type{T}, nn::SparseMatrixCSC, k::Integer, prune::T) where {T<:Union{Float32,Float64}}
    n = size(nn, 2)
    pattern = SparseMatrixCSC(size(nn, 1), n, nn.colptr, nn.rowval, ones(Int32, length(nn.rowval)))   # `trues` graph: pattern only
    d = upload(pattern)
    h = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:svb_jaccard_index, libsvb), Cint, (Ptr{Cvoid}, Int64, Cdouble, Cint, Ref{Ptr{Cvoid}}),
        d, k, Float64(prune), svbtype(T), h))
    o = DeviceMatrix(h[])
    nz = Ref{Int64}(0)
    check(ccall((:svb_matrix_info, libsvb), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ref{Int64}, Ptr{Cint}), o, C_NULL, C_NULL, nz, C_NULL))
    colptr = Vector{Int64}(undef, n + 1); rowval = Vector{Int64}(undef, nz[]); nzval = Vector{T}(undef, nz[])
    check(ccall((:svb_matrix_download, libsvb), Cint, (Ptr{Cvoid}, Ptr{Int64}, Ptr{Int64}, Ptr{Cvoid}, Cint, Cint),
        o, colptr, rowval, nzval, svbtype(T), 1))
    SparseMatrixCSC(n, n, colptr, rowval, nzval)
end
_jaccard_index(::Type{T}, nn::SparseMatrixCSC, prune::T) where {T<:Union{Float32,Float64}} = _jaccard_index(T, nn, 0, prune)

function jaccard_index(nn::NamedArray; prune::Real=1/15, dtype::Type{R}=Float64) where {R<:AbstractFloat}
    snn = _jaccard_index(dtype, nn.array, convert(dtype, prune))
    NamedArray(snn, (nn.dicts[1], nn.dicts[1]), (nn.dimnames[1], nn.dimnames[1]))
end
function jaccard_index(nn::NamedArray, k::Int64; prune::Real=1/15, dtype::Type{R}=Float64) where {R<:AbstractFloat}
    snn = _jaccard_index(dtype, nn.array, k, convert(dtype, prune))
    NamedArray(snn, (nn.dicts[1], nn.dicts[1]), (nn.dimnames[1], nn.dimnames[1]))
end

function shared_nearest_neighbours(rng, X::NamedArray{T,2}, k::Int64; dims=:, metric=Severo.Euclidean(), include_self::Bool=true,
        ntables::Int64=2*size(X,2), prune::Real=1/15) where {T}
    nn = nearest_neighbours(rng, X, k; dims=dims, metric=metric, include_self=include_self, ntables=ntables)
    snn = _jaccard_index(T, nn.array, k, convert(T, prune))                                          # neighbours.jl:267-268
    NamedArray(snn, (X.dicts[1], X.dicts[1]), (X.dimnames[1], X.dimnames[1]))
end
shared_nearest_neighbours(X::NamedArray, k::Int64; kw...) = shared_nearest_neighbours(Random.default_rng(), X, k; kw...)
shared_nearest_neighbours(rng, em::LinearEmbedding, k::Int64; kw...) = shared_nearest_neighbours(rng, em.coordinates, k; kw...)
shared_nearest_neighbours(em::LinearEmbedding, k::Int64; kw...) = shared_nearest_neighbours(Random.default_rng(), em, k; kw...)

# ---- one process, one calling thread, N GPUs (csrc/multi.cu) -------------------------------------------------------------
# The call shape of irlba.jl:66-71 on a whole node: `init_devices()` once, then the WHOLE SparseMatrixCSC goes down in one
# `ccall`; the library shards it by cells over its worker threads (one per GPU) and fills U, s, V. No launcher, no MPI.
init_devices(ndev::Integer=0) = check(ccall((:svb_init_devices, libsvb), Cint, (Cint, Ptr{Cint}), ndev, C_NULL))
shutdown_devices() = check(ccall((:svb_shutdown_devices, libsvb), Cint, ()))

function irlba_devices(A::SparseMatrixCSC{T,Int64}, nu::Integer; mu=nothing, tol::Real=1e-5, maxit::Integer=1000,
                       init::AbstractVector=randn(size(A, 2))) where {T<:Union{Float32,Float64}}
    m, n = size(A)
    U = zeros(m, nu); s = zeros(nu); V = zeros(n, nu)
    iter = Ref{Int64}(0); mprod = Ref{Int64}(0)
    info = ccall((:svb_irlba_csc_devices, libsvb), Cint,
        (Int64, Int64, Ptr{Int64}, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cint, Cint, Ptr{Float64}, Int64, Int64, Int64, Float64, Float64,
         Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Int64}, Ref{Int64}),
        m, n, A.colptr, A.rowval, SVB_I64, A.nzval, svbtype(T), 1, _mu_ptr(mu), nu, min(nu + 7, min(m, n)), maxit, tol, tol,
        convert(Vector{Float64}, init), s, U, V, iter, mprod)
    info == 0 || error("convergence failed")                                   # irlba.jl:73
    SVD(U, s, V')                                                                # irlba.jl:75
end
irlba_devices(C::CenteredMatrix, nu::Integer; kw...) = irlba_devices(C.A isa NamedArray ? C.A.array : C.A, nu; mu=C.mu isa NamedArray ? C.mu.array : C.mu, kw...)

# the fused PCA over the raw counts of the HVG columns on all GPUs: (SVD, stored centre mean/sd)
function pca_counts_devices(counts_hvg::SparseMatrixCSC{T,Int64}, libsize::Vector{Int64}, npcs::Integer; scale_factor::Real=1e4,
                            scale_max::Real=Inf, tol::Real=1e-5, maxit::Integer=1000, init::AbstractVector=randn(size(counts_hvg, 2))) where {T<:Union{Int32,Int64}}
    m, n = size(counts_hvg)
    U = zeros(m, npcs); s = zeros(npcs); V = zeros(n, npcs); mu = zeros(n)
    iter = Ref{Int64}(0); mprod = Ref{Int64}(0)
    info = ccall((:svb_pca_counts_devices, libsvb), Cint,
        (Int64, Int64, Ptr{Int64}, Ptr{Cvoid}, Cint, Ptr{Cvoid}, Cint, Cint, Ptr{Int64}, Float64, Float64, Int64, Int64, Int64, Float64,
         Float64, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ptr{Float64}, Ref{Int64}, Ref{Int64}),
        m, n, counts_hvg.colptr, counts_hvg.rowval, SVB_I64, counts_hvg.nzval, svbtype(T), 1, libsize, scale_factor, scale_max, npcs,
        min(npcs + 7, min(m, n)), maxit, tol, tol, convert(Vector{Float64}, init), mu, s, U, V, iter, mprod)
    info == 0 || error("convergence failed")
    SVD(U, s, V'), mu
end

end # module
