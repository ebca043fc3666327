# CEGCuda.jl — the reference-side binding of libceg_cuda (include/ceg_cuda.h).
#
# Drop this file next to CrystalEnergyGrids.jl's src/ and `include("CEGCuda.jl")` at the end of
# src/CrystalEnergyGrids.jl (after grids.jl, probes.jl and ewald.jl).  It re-defines, with identical
# signatures, the entry points of the hot path so that their inner loops run on the GPU:
#
#   create_grid_vdw        grids.jl:137-157      -> ceg_probe_create + ceg_grid_build
#   create_grid_coulomb    grids.jl:159-185      -> ceg_probe_create + ceg_grid_build
#   interpolate_grid       grids.jl:212-273      -> ceg_grid_upload (once per grid) + ceg_interpolate
#   energy_point(setup, positions, ctx)   grids.jl:311-327   -> ceg_energy_points   (METHOD OVERRIDE, same signature)
#   energy_grid(setup, step, num_rotate)  grids.jl:346-424   -> ceg_energy_lattice  (METHOD OVERRIDE for CrystalEnergySetup)
#   compute_ewald(ctx, skip)              ewald.jl:555-577   -> ceg_ewald_reciprocal (METHOD OVERRIDE)
#   initialize_ewald                      ewald.jl:195-281   -> unchanged host code; its EwaldFramework gets a device handle on
#                                                               first use (ceg_ewald_create_from_sf), or — with the three-line patch
#                                                               of INTEGRATION.md §2 — S_f itself comes from ceg_ewald_create
#   setup_RASPA's grid hand-off           raspa.jl:420-439   -> create_grids_resident (ceg_grid_create_empty + ceg_grid_build_records:
#                                                               the build kernel writes the interpolation records itself)
#   multi-GPU, one process per GPU         raspa.jl:430-438   -> ShardedGridBuild / build! (ceg_ipc_export/open + ceg_grid_build_records:
#                                                               every rank's kernel writes its slab into every rank's grids)
#   single_contribution_ewald  ewald.jl:704-738  -> ceg_ewald_single_contribution   (batched over trial placements)
#   single_contribution_vdw    energy.jl:397-427 -> ceg_guest_energy                (batched over trial placements)
#   parse_blockfile / BlockFile(g)  coordinates.jl:112-167, grids.jl:188-204 -> ceg_block_from_spheres / _values
#
# Everything else (force-field parsing, CIF loading, .grid headers, caches, Monte-Carlo) is
# untouched Julia.  Units are stripped to Å / K / e before each `ccall`; every physical constant is
# passed as an argument, so the Unitful version in use keeps deciding GRID_TO_KELVIN and
# COULOMBIC_CONVERSION_FACTOR.
#
# NOTE: Julia is not installed in the build image of this repository, so this file has not been
# executed there; the same C ABI is exercised through the ctypes mirror
# (crystalenergygrids.jl_b200/_lib.py) by the GPU test-suite.

module CEGCuda

using ..CrystalEnergyGrids
using ..CrystalEnergyGrids: ForceField, InteractionRule, InteractionRuleSum, FF, ProbeSystem, EnergyGrid,
    GridCoordinatesSetup, EwaldFramework, EwaldContext, CellMatrix, BlockFile, CrystalEnergySetup,
    prepare_periodic_distance_computations, find_supercell, initialize_ewald, _setup_grid_common,
    _create_grid_common, GRID_TO_KELVIN, COULOMBIC_CONVERSION_FACTOR, TÅ, TK, isdefined_ewald, get_rotation_matrices,
    ChangePositionSystem, bounding_box, hash_string, lebedev_num, scratchspace
using LinearAlgebra: norm
using Serialization: serialize, deserialize
using StaticArrays, Unitful, UnitfulAtomic

const libceg = get(ENV, "LIBCEG_CUDA", "libceg_cuda.so")

# ---- C structs (include/ceg_cuda.h) -----------------------------------------------------------------
struct CegCell
    mat::NTuple{9,Cdouble}
    invmat::NTuple{9,Cdouble}
    ortho::Int32
    _pad::Int32
    safemin2::Cdouble
end
struct CegGridSpec
    dims::NTuple{3,Int32}
    _pad::Int32
    size::NTuple{3,Cdouble}
    shift::NTuple{3,Cdouble}
    delta::NTuple{3,Cdouble}
end
struct CegPair
    flags::Int32
    _pad::Int32
    hs_r::Cdouble
    lj_eps::Cdouble; lj_sigma::Cdouble; lj_shift::Cdouble
    buck_A::Cdouble; buck_B::Cdouble; buck_C::Cdouble; buck_shift::Cdouble
end

const PAIR_HS, PAIR_LJ, PAIR_BUCK = Int32(1), Int32(2), Int32(4)

ceg_error() = unsafe_string(ccall((:ceg_last_error, libceg), Cstring, ()))
check(rc::Cint) = rc == 0 ? nothing : error(ceg_error())   # same failure mode as the reference: `error(...)`

function CegCell(mat::SMatrix{3,3,Float64,9}, invmat::SMatrix{3,3,Float64,9})
    _, ortho, safemin = prepare_periodic_distance_computations(mat)          # utils.jl:146-155
    CegCell(Tuple(mat), Tuple(invmat), Int32(ortho), Int32(0), safemin^2)
end
function CegGridSpec(cs::GridCoordinatesSetup)
    å(x) = Tuple(Float64.(ustrip.(u"Å", x)))
    CegGridSpec(Tuple(Int32.(cs.dims)), Int32(0), å(cs.size), å(cs.shift), å(cs.Δ))
end

# derivativesGrid's view of one ForceField entry (interactions.jl:432-472, 599-610)
function CegPair(rule::Union{InteractionRule,InteractionRuleSum})
    flags = Int32(0); hs = 0.0; lj = (0.0, 0.0, 0.0); bk = (0.0, 0.0, 0.0, 0.0)
    for r in (rule isa InteractionRule ? (rule,) : rule.rules)
        k = r.kind
        if k === FF.NoInteraction || k === FF.CoulombEwaldDirect
            continue
        elseif k === FF.HardSphere
            flags |= PAIR_HS; hs = r.params[1] + r.params[2]
        elseif k === FF.LennardJones
            flags |= PAIR_LJ; lj = (r.params[1], r.params[2], r.shift)
        elseif k === FF.Buckingham
            flags |= PAIR_BUCK; bk = (r.params[1], r.params[2], r.params[3], r.shift)
        elseif k === FF.UndefinedInteraction
            flags = Int32(-1)
        else
            error("VdW grid not implemented for $(k)")                        # interactions.jl:442,462-465
        end
    end
    CegPair(flags, 0, hs, lj..., bk...)
end

# ---- grid construction ---------------------------------------------------------------------------------
mutable struct ProbeHandle
    ptr::Ptr{Cvoid}
    function ProbeHandle(s::ProbeSystem, probes::Vector{Int}, coulomb::Bool, α::Float64; device=0)
        ff = s.forcefield
        ntypes = length(ff.sdict)
        table = [CegPair(ff.interactions[t, p]) for p in probes for t in 1:ntypes]   # row p, column t
        cell = CegCell(s.mat, s.invmat)
        xyz = reinterpret(Float64, s.positions)                                      # n x 3, row-major
        types = Int32.(s.atomkinds .- 1)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        GC.@preserve xyz types table begin
            check(ccall((:ceg_probe_create, libceg), Cint,
                (Int32, Ref{CegCell}, Int64, Ptr{Cdouble}, Ptr{Int32}, Ptr{Cdouble}, Int32, Int32, Ptr{CegPair},
                 Cdouble, Int32, Cdouble, Ref{Ptr{Cvoid}}),
                device, cell, length(s.positions), xyz, types, coulomb ? s.charges : C_NULL, ntypes,
                length(probes), table, ustrip(u"Å^2", ff.cutoff^2), coulomb, α, out))
        end
        h = new(out[])
        finalizer(x -> ccall((:ceg_probe_free, libceg), Cint, (Ptr{Cvoid},), x.ptr), h)
    end
end

function _build!(grids::Vector{Array{Cfloat,4}}, h::ProbeHandle, cset::GridCoordinatesSetup;
                 xrange=0:Int(cset.dims[1]))
    λv = inv(GRID_TO_KELVIN); Tv = GRID_TO_KELVIN * 1e7                                # grids.jl:142-148
    λc = ustrip(u"K*Å/e_au^2", COULOMBIC_CONVERSION_FACTOR) / GRID_TO_KELVIN; Tc = inv(λc) * 1e7   # :169-170
    ptrs = [pointer(g) for g in grids]
    GC.@preserve grids begin
        check(ccall((:ceg_grid_build, libceg), Cint,
            (Ptr{Cvoid}, Ref{CegGridSpec}, Int32, Int32, Cdouble, Cdouble, Cdouble, Cdouble, Ptr{Ptr{Cfloat}}, Ptr{Cdouble}),
            h.ptr, CegGridSpec(cset), first(xrange), last(xrange) + 1, λv, Tv, λc, Tc, ptrs, C_NULL))
    end
    grids
end

# grids.jl:137-157 — same signature, same file layout, same return value
function CrystalEnergyGrids.create_grid_vdw(file, framework, forcefield::ForceField, spacing::TÅ, atom::Symbol)
    cset, num_unitcell = _setup_grid_common(framework, spacing, forcefield.cutoff)
    grid = Array{Cfloat,4}(undef, cset.dims[3]+1, cset.dims[2]+1, cset.dims[1]+1, 8)
    probe = ProbeSystem(framework, forcefield, atom)
    _build!([grid], ProbeHandle(probe, [probe.probe], false, 0.0), cset)
    open(file, "w") do f
        _create_grid_common(f, cset, num_unitcell)
        write(f, grid)
        write(f, NoUnits.(cset.cell.mat ./ u"Å"))
    end
    grid
end

# grids.jl:159-185
function CrystalEnergyGrids.create_grid_coulomb(file, framework, forcefield::ForceField, spacing::TÅ, _ewald=nothing)
    cset, num_unitcell = _setup_grid_common(framework, spacing, 12.0u"Å")
    ewald = _ewald isa EwaldFramework ? _ewald : initialize_ewald(framework, num_unitcell)
    grid = Array{Cfloat,4}(undef, cset.dims[3]+1, cset.dims[2]+1, cset.dims[1]+1, 8)
    probe = ProbeSystem(framework, forcefield)
    _build!([grid], ProbeHandle(probe, Int[], true, ustrip(u"Å^-1", ewald.α)), cset)
    open(file, "w") do f
        _create_grid_common(f, cset, num_unitcell)
        write(f, ewald.precision)
        write(f, grid)
        write(f, NoUnits.(cset.cell.mat ./ u"Å"))
    end
    grid
end

# New: all VdW probes of a molecule and the Coulomb grid in ONE pass over the framework atoms
# (what retrieve_or_create_grid's callers, raspa.jl:499-516, loop over one grid at a time).
function create_grids_fused(framework, forcefield::ForceField, spacing::TÅ, atoms::Vector{Symbol}, ewald::Union{Nothing,EwaldFramework})
    cset, _ = _setup_grid_common(framework, spacing, forcefield.cutoff)
    probe = ProbeSystem(framework, forcefield)                                   # positions/charges/kinds
    idx = [forcefield.sdict[Symbol(CrystalEnergyGrids.get_atom_name(a))] for a in atoms]
    n = length(idx) + !(ewald isa Nothing)
    grids = [Array{Cfloat,4}(undef, cset.dims[3]+1, cset.dims[2]+1, cset.dims[1]+1, 8) for _ in 1:n]
    α = ewald isa Nothing ? 0.0 : ustrip(u"Å^-1", ewald.α)
    _build!(grids, ProbeHandle(probe, idx, !(ewald isa Nothing), α), cset)
end

# ---- interpolation -----------------------------------------------------------------------------------------
const GRID_HANDLES = IdDict{Array{Cfloat,4},Ptr{Cvoid}}()      # device copy per EnergyGrid.grid
const GRID_LOCK = ReentrantLock()

function handle(g::EnergyGrid; device=0)
    lock(GRID_LOCK) do
        get!(GRID_HANDLES, g.grid) do
            cell = CegCell(SMatrix{3,3,Float64,9}(ustrip.(u"Å", g.csetup.cell.mat)),
                           SMatrix{3,3,Float64,9}(ustrip.(u"Å^-1", g.csetup.cell.invmat)))
            out = Ref{Ptr{Cvoid}}(C_NULL)
            check(ccall((:ceg_grid_upload, libceg), Cint,
                (Int32, Ref{CegCell}, Ref{CegGridSpec}, Ptr{Cfloat}, Int32, Ref{Ptr{Cvoid}}),
                device, cell, CegGridSpec(g.csetup), g.grid, g.ewald_precision == Inf, out))
            out[]
        end
    end
end

# A grid that is already on the device in `.grid` order and units (ceg_grid_build_dev output, or slabs gathered from
# other GPUs): the handle is made on the device with parse_grid's conversion (grids.jl:78), no file, no host copy.
function handle_from_device(cs::GridCoordinatesSetup, d_grid::Ptr{Cfloat}, is_vdw::Bool; device=0, scale=GRID_TO_KELVIN)
    cell = CegCell(SMatrix{3,3,Float64,9}(ustrip.(u"Å", cs.cell.mat)), SMatrix{3,3,Float64,9}(ustrip.(u"Å^-1", cs.cell.invmat)))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ceg_grid_upload_dev, libceg), Cint,
        (Int32, Ref{CegCell}, Ref{CegGridSpec}, Ptr{Cfloat}, Cdouble, Int32, Ref{Ptr{Cvoid}}),
        device, cell, CegGridSpec(cs), d_grid, scale, is_vdw, out))
    out[]
end

# batched method of interpolate_grid (grids.jl:212-273): points is 3 x n (Å); returns K, 1e100 = blocked
function CrystalEnergyGrids.interpolate_grid(g::EnergyGrid, points::AbstractMatrix{Float64})
    g.ewald_precision === -Inf32 && return zeros(size(points, 2))
    isnan(g.ewald_precision) && error("Invalid grid cannot be interpolated!")
    out = Vector{Float64}(undef, size(points, 2))
    check(ccall((:ceg_interpolate, libceg), Cint, (Ptr{Cvoid}, Int64, Ptr{Cdouble}, Ptr{Cdouble}),
                handle(g), size(points, 2), points, out))
    out
end

# ---- Ewald ---------------------------------------------------------------------------------------------------
# S_f(k) of initialize_ewald (ewald.jl:264-271): call with the k-list/kfactors computed by the Julia code.
function structure_factor(invmat::SMatrix{3,3,Float64,9}, ks::NTuple{3,Int}, kindices, kfactors::Vector{Float64},
                          frac::Matrix{Float64}, charges::Vector{Float64}; device=0)
    kvec = Int32[]
    for (jy, jz, jxrange, _) in kindices, ix in jxrange
        append!(kvec, (ix, jy, jz))
    end
    S = Matrix{Float64}(undef, 2, length(kfactors))
    out = Ref{Ptr{Cvoid}}(C_NULL)
    check(ccall((:ceg_ewald_create, libceg), Cint,
        (Int32, Ptr{Cdouble}, Ptr{Int32}, Int64, Ptr{Int32}, Ptr{Cdouble}, Int64, Ptr{Cdouble}, Ptr{Cdouble},
         Ref{Ptr{Cvoid}}, Ptr{Cdouble}),
        device, collect(invmat), Int32[ks...], length(kfactors), kvec, kfactors, length(charges), frac, charges, out, S))
    out[], reinterpret(ComplexF64, vec(S))
end

# compute_ewald(ctx) for many placements of the context's molecule (ewald.jl:350-364, 555-577)
function reciprocal(h::Ptr{Cvoid}, positions::Array{Float64,3}, q::Vector{Float64}, energy_net_charges, static_contribution)
    n_atoms, n_pos = size(positions, 2), size(positions, 3)
    out = Vector{Float64}(undef, n_pos)
    check(ccall((:ceg_ewald_reciprocal, libceg), Cint,
        (Ptr{Cvoid}, Int64, Int32, Ptr{Cdouble}, Ptr{Cdouble}, Cdouble, Cdouble, Ptr{Cdouble}),
        h, n_pos, n_atoms, positions, q, ustrip(u"K", energy_net_charges), ustrip(u"K", static_contribution), out))
    out
end

# single_contribution_ewald (ewald.jl:704-738) for many trial placements of molecule ij:
# rest = ewald.eframework.StoreRigidChargeFramework + ewald.sums[:,1] - ewald.sums[:,ij+1]
function single_contributions(h::Ptr{Cvoid}, rest::Vector{ComplexF64}, positions::Array{Float64,3}, q::Vector{Float64})
    n_atoms, n_pos = size(positions, 2), size(positions, 3)
    out = Vector{Float64}(undef, n_pos)
    check(ccall((:ceg_ewald_single_contribution, libceg), Cint,
        (Ptr{Cvoid}, Ptr{Cdouble}, Int64, Int32, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{Cdouble}),
        h, reinterpret(Float64, rest), n_pos, n_atoms, positions, q, out))
    out
end

# ---- batched energy_point (grids.jl:311-327; inner loop of energy_grid, :394-419) ------------------------------
# positions: 3 x n_atoms x n_pos (Å).  vdw_handles[i] may be C_NULL (zero grid); coulomb / ewald / block may be C_NULL.
function energy_points(vdw_handles::Vector{Ptr{Cvoid}}, coulomb::Ptr{Cvoid}, ewald::Ptr{Cvoid}, block::Ptr{Cvoid},
                       atomsidx::Vector{Int}, charges::Vector{Float64}, positions::Array{Float64,3},
                       energy_net_charges, static_contribution)
    n_atoms, n_pos = size(positions, 2), size(positions, 3)
    vdw = Vector{Float64}(undef, n_pos); coul = Vector{Float64}(undef, n_pos)
    check(ccall((:ceg_energy_points, libceg), Cint,
        (Int32, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Int64, Int32, Ptr{Int32}, Ptr{Cdouble}, Ptr{Cdouble},
         Cdouble, Cdouble, Ptr{Cdouble}, Ptr{Cdouble}),
        length(vdw_handles), vdw_handles, coulomb, ewald, block, n_pos, n_atoms, Int32.(atomsidx .- 1), charges, positions,
        ustrip(u"K", energy_net_charges), ustrip(u"K", static_contribution), vdw, coul))
    vdw, coul
end

# energy_grid (grids.jl:346-424): the lattice (iA-1)stepA + (iB-1)stepB + (iC-1)stepC + rotpos is generated on the
# device.  steps: 3 x 3 (column d = axis_d / num_d, Å); rotpos: 3 x n_atoms x n_rot.  Returns (numC, numB, numA, n_rot)
# in Julia's column-major order, i.e. permutedims(., (4,3,2,1)) of the reference's allvals.
function energy_lattice(vdw_handles::Vector{Ptr{Cvoid}}, coulomb::Ptr{Cvoid}, ewald::Ptr{Cvoid}, block::Ptr{Cvoid},
                        nums::NTuple{3,Int}, steps::Matrix{Float64}, rotpos::Array{Float64,3}, atomsidx::Vector{Int},
                        charges::Vector{Float64}, energy_net_charges, static_contribution)
    out = Array{Float64,4}(undef, nums[3], nums[2], nums[1], size(rotpos, 3))
    check(ccall((:ceg_energy_lattice, libceg), Cint,
        (Int32, Ptr{Ptr{Cvoid}}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Cvoid}, Ptr{Int32}, Ptr{Cdouble}, Int32, Int32, Ptr{Cdouble},
         Ptr{Int32}, Ptr{Cdouble}, Cdouble, Cdouble, Ptr{Cdouble}),
        length(vdw_handles), vdw_handles, coulomb, ewald, block, Int32[nums...], steps, size(rotpos, 3), size(rotpos, 2), rotpos,
        Int32.(atomsidx .- 1), charges, ustrip(u"K", energy_net_charges), ustrip(u"K", static_contribution), out))
    out
end

# ---- blocking masks ------------------------------------------------------------------------------------------------
# BlockFile(g) (grids.jl:188-204): mask[c, b, a] (Julia order of the (a,b,c) BitArray transposed to C order)
function block_from_values(cs::GridCoordinatesSetup, values::Array{Cfloat,3}; threshold=5f6, device=0)
    mask = Vector{UInt8}(undef, length(values))
    check(ccall((:ceg_block_from_values, libceg), Cint, (Int32, Ref{CegGridSpec}, Ptr{Cfloat}, Cfloat, Ptr{UInt8}),
                device, CegGridSpec(cs), values, threshold, mask))
    mask
end
# parse_blockfile (coordinates.jl:112-167): centers 3 x n (Å, already through inverse_offsetpoint(offsetpoint(.)))
function block_from_spheres(cell::CegCell, cs::GridCoordinatesSetup, centers::Matrix{Float64}, radius2::Vector{Float64}; device=0)
    mask = Vector{UInt8}(undef, prod(Int.(cs.dims) .+ 1))
    check(ccall((:ceg_block_from_spheres, libceg), Cint,
        (Int32, Ref{CegCell}, Ref{CegGridSpec}, Int32, Ptr{Cdouble}, Ptr{Cdouble}, Ptr{UInt8}),
        device, cell, CegGridSpec(cs), length(radius2), centers, radius2, mask))
    mask
end

# ---- guest-guest pair energies (energy.jl:397-427) ---------------------------------------------------------------------
struct CegPairEnergy
    flags::Int32
    _pad::Int32
    hs_r::Cdouble
    lj_eps::Cdouble; lj_sigma::Cdouble; lj_shift::Cdouble
    buck_A::Cdouble; buck_B::Cdouble; buck_C::Cdouble; buck_shift::Cdouble
    ewald_alpha::Cdouble; ewald_qq::Cdouble; coulomb_qq::Cdouble
end
# trial_xyz: 3 x n_atoms x n_trials; env_xyz: 3 x n_env; species / molecule ids 0-based Int32
function guest_energies(cell::CegCell, cutoff2::Float64, table::Matrix{CegPairEnergy}, env_xyz::Matrix{Float64},
                        env_species::Vector{Int32}, env_mol::Vector{Int32}, mol_species::Vector{Int32}, skip_mol::Integer,
                        trial_xyz::Array{Float64,3}; device=0)
    out = Vector{Float64}(undef, size(trial_xyz, 3))
    check(ccall((:ceg_guest_energy, libceg), Cint,
        (Int32, Ref{CegCell}, Cdouble, Int32, Ptr{CegPairEnergy}, Int64, Ptr{Cdouble}, Ptr{Int32}, Ptr{Int32}, Int64, Int32,
         Ptr{Int32}, Int32, Ptr{Cdouble}, Ptr{Cdouble}),
        device, cell, cutoff2, size(table, 1), permutedims(table), size(env_xyz, 2), env_xyz, env_species, env_mol,
        size(trial_xyz, 3), size(trial_xyz, 2), mol_species, skip_mol, trial_xyz, out))
    out
end

# The environment resident on the device: create once per simulation, evaluate batches, move atoms after accepted steps.
mutable struct GuestEnv
    ptr::Ptr{Cvoid}
    function GuestEnv(cell::CegCell, cutoff2::Float64, table::Matrix{CegPairEnergy}, env_xyz::Matrix{Float64},
                      env_species::Vector{Int32}, env_mol::Vector{Int32}; device=0)
        out = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:ceg_guest_env_create, libceg), Cint,
            (Int32, Ref{CegCell}, Cdouble, Int32, Ptr{CegPairEnergy}, Int64, Ptr{Cdouble}, Ptr{Int32}, Ptr{Int32}, Ref{Ptr{Cvoid}}),
            device, cell, cutoff2, size(table, 1), permutedims(table), size(env_xyz, 2), env_xyz, env_species, env_mol, out))
        h = new(out[])
        finalizer(x -> ccall((:ceg_guest_env_free, libceg), Cint, (Ptr{Cvoid},), x.ptr), h)
    end
end
function guest_energies(env::GuestEnv, mol_species::Vector{Int32}, skip_mol::Integer, trial_xyz::Array{Float64,3})
    out = Vector{Float64}(undef, size(trial_xyz, 3))
    check(ccall((:ceg_guest_env_energy, libceg), Cint, (Ptr{Cvoid}, Int64, Int32, Ptr{Int32}, Int32, Ptr{Cdouble}, Ptr{Cdouble}),
                env.ptr, size(trial_xyz, 3), size(trial_xyz, 2), mol_species, skip_mol, trial_xyz, out))
    out
end
move!(env::GuestEnv, first::Integer, xyz::Matrix{Float64}) =      # atoms first..first+n-1 (1-based) take the columns of xyz
    check(ccall((:ceg_guest_env_update, libceg), Cint, (Ptr{Cvoid}, Int64, Int64, Ptr{Cdouble}), env.ptr, first - 1, size(xyz, 2), xyz))

# ---- interpolation sweep over device-generated positions (benchmarking / sampling) -------------------------------------
function interpolate_sweep(g::EnergyGrid, n::Integer, seed::Integer; first=0, n_keep=0)
    mat = collect(Float64.(ustrip.(u"Å", g.csetup.cell.mat)))            # column-major 3x3
    s = Ref(0.0); mn = Ref(0.0); nb = Ref{Int64}(0); ms = Ref{Cfloat}(0)
    keep = Vector{Float64}(undef, min(n, n_keep))
    check(ccall((:ceg_interpolate_sweep, libceg), Cint,
        (Ptr{Cvoid}, Int64, Int64, UInt64, Ptr{Cdouble}, Ref{Cdouble}, Ref{Cdouble}, Ref{Int64}, Int64, Ptr{Cdouble}, Ref{Cfloat}),
        handle(g), first, n, seed, mat, s, mn, nb, length(keep), keep, ms))
    (sum=s[], min=mn[], n_blocked=nb[], keep=keep, kernel_ms=ms[])
end

# =====================================================================================================================
# Method overrides with the reference's own signatures: after `include("CEGCuda.jl")` the calls below run on the GPU.
# =====================================================================================================================

# ---- device handles of the Julia objects (created on first use, keyed by object identity) ---------------------------------
const EWALD_HANDLES = IdDict{Vector{ComplexF64},Ptr{Cvoid}}()       # EwaldFramework.StoreRigidChargeFramework -> ceg_ewald_h
const BLOCK_HANDLES = IdDict{Any,Ptr{Cvoid}}()                      # BlockFile.block -> ceg_block_h
const HANDLE_LOCK = ReentrantLock()

function _kvec(kspace)
    kvec = Int32[]
    for (jy, jz, jxrange, _) in kspace.kindices, ix in jxrange
        append!(kvec, (ix, jy, jz))
    end
    kvec
end

# EwaldFramework of an unchanged initialize_ewald (ewald.jl:195-281) -> device handle; nothing is re-summed
function handle(ef::EwaldFramework; device=0)
    iszero(ef.α) && return C_NULL
    lock(HANDLE_LOCK) do
        get!(EWALD_HANDLES, ef.StoreRigidChargeFramework) do
            out = Ref{Ptr{Cvoid}}(C_NULL)
            invmat = collect(Float64.(ustrip.(u"Å^-1", ef.invmat)))
            check(ccall((:ceg_ewald_create_from_sf, libceg), Cint,
                (Int32, Ptr{Cdouble}, Ptr{Int32}, Int64, Ptr{Int32}, Ptr{Cdouble}, Ptr{Cdouble}, Ref{Ptr{Cvoid}}),
                device, invmat, Int32[ef.kspace.ks...], ef.kspace.num_kvecs, _kvec(ef.kspace), ef.kfactors,
                reinterpret(Float64, ef.StoreRigidChargeFramework), out))
            out[]
        end
    end
end

# BlockFile (coordinates.jl:83-101) -> device mask; C_NULL for an empty one
function handle(b::BlockFile; device=0)
    b.empty && return C_NULL
    lock(HANDLE_LOCK) do
        get!(BLOCK_HANDLES, b.block) do
            cs = b.csetup
            cell = CegCell(SMatrix{3,3,Float64,9}(ustrip.(u"Å", cs.cell.mat)), SMatrix{3,3,Float64,9}(ustrip.(u"Å^-1", cs.cell.invmat)))
            mask = UInt8.(permutedims(b.block, (3, 2, 1)))             # (a,b,c) BitArray -> [a][b][c] bytes
            out = Ref{Ptr{Cvoid}}(C_NULL)
            check(ccall((:ceg_block_upload, libceg), Cint, (Int32, Ref{CegCell}, Ref{CegGridSpec}, Ptr{UInt8}, Ref{Ptr{Cvoid}}),
                        device, cell, CegGridSpec(cs), mask, out))
            out[]
        end
    end
end

_gridhandle(g::EnergyGrid) = (g.ewald_precision == -Inf || isempty(g.grid)) ? C_NULL : handle(g)

# static part of compute_ewald for ONE molecule of the setup (ewald.jl:484-531): energy_net_charges, static_contribution
function _ewald_constants(setup::CrystalEnergySetup)
    ctx = EwaldContext(setup.ewald, ((setup.molecule,),))
    isdefined_ewald(ctx) || return 0.0, 0.0
    ustrip(u"K", ctx.energy_net_charges), ustrip(u"K", ctx.static_contribution[])
end

function _handles(setup::CrystalEnergySetup)
    for g in setup.grids
        isnan(g.ewald_precision) && error("Invalid grid cannot be interpolated!")
    end
    vdw = Ptr{Cvoid}[_gridhandle(g) for g in setup.grids]
    hascoulomb = setup.coulomb.ewald_precision != -Inf
    coul = hascoulomb ? _gridhandle(setup.coulomb) : C_NULL
    ew = hascoulomb ? handle(setup.ewald) : C_NULL
    vdw, coul, ew, handle(setup.block)
end

"""
    energy_point(setup::CrystalEnergySetup, positions, ctx=nothing)          # grids.jl:311-327, same contract

One placement through the library's latency path (a single kernel launch).  `ctx` is accepted for signature
compatibility; the reciprocal term is evaluated from `positions` on the device (equal to
`move_one_system!(ctx, 1, positions); compute_ewald(ctx)`), so `ctx` is left untouched.
"""
function CrystalEnergyGrids.energy_point(setup::CrystalEnergySetup, positions, ctx=nothing)
    n = length(setup.atomsidx)
    pos = Array{Float64,3}(undef, 3, n, 1)
    for i in 1:n, c in 1:3
        pos[c, i, 1] = ustrip(u"Å", positions[i][c])
    end
    vdw, coul, ew, blk = _handles(setup)
    enc, stat = _ewald_constants(setup)
    v, c = energy_points(vdw, coul, ew, blk, setup.atomsidx, setup.charges, pos, enc * u"K", stat * u"K")
    v[1] == 1e100 && return (1e100u"K", 0.0u"K")
    (v[1] * u"K", c[1] * u"K")
end

"""
    energy_grid(setup::CrystalEnergySetup, step, num_rotate=40)             # grids.jl:346-424, same contract

Same lattice ((iA-1)stepA + (iB-1)stepB + (iC-1)stepC, left-to-right sum), same Lebedev orientations
(`get_rotation_matrices`, unchanged lebedev.jl), same scratch-space cache; the numA x numB x numC x n_rot
`energy_point` evaluations run as one batched call with the lattice generated on the device.
(`num_rotate < 0`, random offsets, and MonteCarloSetup inputs keep the original Julia method.)
"""
function CrystalEnergyGrids.energy_grid(setup::CrystalEnergySetup, step, num_rotate=40)
    num_rotate < 0 && return invoke(CrystalEnergyGrids.energy_grid, Tuple{Any,Any,Any}, setup, step, num_rotate)
    molecule = setup.molecule
    store = string(hash_string(setup), '-', ustrip(u"Å", step), '-', lebedev_num(molecule, num_rotate))
    path = joinpath(scratchspace, store)
    isfile(path) && return deserialize(path)::Array{Float64,4}
    __pos = position(molecule) / u"Å"
    axes = bounding_box(setup.framework)
    nums = ntuple(d -> floor(Int, norm(axes[d]) / step) + 1, 3)
    steps = Matrix{Float64}(undef, 3, 3)                       # column d = axis_d / num_d (Å)
    for d in 1:3
        steps[:, d] .= ustrip.(u"Å", axes[d] ./ nums[d])
    end
    rots = (num_rotate == 0 || length(molecule) == 1) ? [one(SMatrix{3,3,Float64,9})] : first(get_rotation_matrices(molecule, num_rotate))
    rotpos = Array{Float64,3}(undef, 3, length(__pos), length(rots))
    for (k, r) in enumerate(rots), (i, p) in enumerate(__pos)
        rotpos[:, i, k] .= r * SVector{3,Float64}(p)
    end
    vdw, coul, ew, blk = _handles(setup)
    enc, stat = _ewald_constants(setup)
    raw = energy_lattice(vdw, coul, ew, blk, nums, steps, rotpos, setup.atomsidx, setup.charges, enc * u"K", stat * u"K")
    allvals = permutedims(raw, (4, 3, 2, 1))                   # (numC,numB,numA,n_rot) -> (n_rot,numA,numB,numC)
    serialize(path, allvals)
    allvals
end

"""
    compute_ewald(ctx::EwaldContext, skipcontribution=0)                     # ewald.jl:555-577, same contract

The positions of the context's charges are read back from its phase tables (`Eikx[2,l] = e^{2πi f_x}`,
`Eiky[ky+2,l]`, `Eikz[kz+2,l]`, ewald.jl:69-86,350-364): S_a(k) only depends on the fractional coordinates modulo 1.
"""
function CrystalEnergyGrids.compute_ewald(ctx::EwaldContext, skipcontribution=0)
    isdefined_ewald(ctx) || return 0.0u"K"
    Eikx, Eiky, Eikz = ctx.Eiks
    _, ky, kz = ctx.eframework.kspace.ks
    keep = [l for (l, (ij, _)) in enumerate(ctx.atoms) if ctx.speciesof[ij] != skipcontribution]
    isempty(keep) && return 2 * ctx.energy_net_charges + ctx.static_contribution[]
    mat = SMatrix{3,3,Float64,9}(ustrip.(u"Å", ctx.eframework.mat))
    pos = Array{Float64,3}(undef, 3, length(keep), 1)
    q = Vector{Float64}(undef, length(keep))
    for (m, l) in enumerate(keep)
        f = SVector{3,Float64}(angle(Eikx[2, l]), angle(Eiky[ky+2, l]), angle(Eikz[kz+2, l])) ./ (2π)
        pos[:, m, 1] .= mat * f
        ij, k = ctx.atoms[l]
        q[m] = ustrip(u"e_au", ctx.allcharges[ctx.speciesof[ij]][k])
    end
    reciprocal(handle(ctx.eframework), pos, q, ctx.energy_net_charges, ctx.static_contribution[])[1] * u"K"
end

# ---- setup_RASPA's hand-off without the file round trip (raspa.jl:420-439) ---------------------------------------------
# Build the missing grids of a setup in one fused pass, keep them on the device, return EnergyGrid objects whose handles
# are already resident (ceg_grid_build_records: the build kernel applies parse_grid's conversion, grids.jl:78, and writes
# the interpolation records itself) and
# write the `.grid` files from one downloaded copy on a background task.  `retrieve_or_create_grid` then only needs
#     haskey(RESIDENT, grid_path) && return RESIDENT[grid_path]
# as its first line.
const RESIDENT = Dict{String,EnergyGrid}()

function create_grids_resident(framework, forcefield::ForceField, spacing::TÅ, atoms::Vector{Symbol}, paths::Vector{String},
                               ewald::Union{Nothing,EwaldFramework}, coulomb_path::String=""; device=0)
    cset, num_unitcell = _setup_grid_common(framework, spacing, forcefield.cutoff)
    probe = ProbeSystem(framework, forcefield)
    idx = [forcefield.sdict[Symbol(CrystalEnergyGrids.get_atom_name(a))] for a in atoms]
    withcoul = !(ewald isa Nothing)
    n = length(idx) + withcoul
    nz, ny, nx = cset.dims[3] + 1, cset.dims[2] + 1, cset.dims[1] + 1
    bytes = 8 * nx * ny * nz * sizeof(Cfloat)
    h = ProbeHandle(probe, idx, withcoul, withcoul ? ustrip(u"Å^-1", ewald.α) : 0.0; device)
    cell = CegCell(SMatrix{3,3,Float64,9}(ustrip.(u"Å", cset.cell.mat)), SMatrix{3,3,Float64,9}(ustrip.(u"Å^-1", cset.cell.invmat)))
    gs = CegGridSpec(cset)
    # empty interpolation grids: the build kernel writes their node records itself (ceg_grid_build_records) ...
    grids = [Ref{Ptr{Cvoid}}(C_NULL) for _ in 1:n]
    recs = Vector{Ptr{Cvoid}}(undef, n)
    layout = Ref{Int32}(0)
    for g in 1:n
        check(ccall((:ceg_grid_create_empty, libceg), Cint, (Int32, Ref{CegCell}, Ref{CegGridSpec}, Int32, Ref{Ptr{Cvoid}}),
            device, cell, gs, !(withcoul && g == n), grids[g]))
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:ceg_grid_records, libceg), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}, Ptr{Int64}, Ref{Int32}), grids[g][], r, C_NULL, layout))
        recs[g] = r[]
    end
    # ... and the `.grid`-order copy for the files on the side
    d = [Ref{Ptr{Cvoid}}(C_NULL) for _ in 1:n]
    foreach(r -> check(ccall((:ceg_device_malloc, libceg), Cint, (Int32, Int64, Ref{Ptr{Cvoid}}), device, bytes, r)), d)
    ptrs = Ptr{Cfloat}[Ptr{Cfloat}(r[]) for r in d]
    λv = inv(GRID_TO_KELVIN); Tv = GRID_TO_KELVIN * 1e7
    λc = ustrip(u"K*Å/e_au^2", COULOMBIC_CONVERSION_FACTOR) / GRID_TO_KELVIN; Tc = inv(λc) * 1e7
    check(ccall((:ceg_grid_build_records, libceg), Cint,
        (Ptr{Cvoid}, Ref{CegGridSpec}, Int32, Int32, Cdouble, Cdouble, Cdouble, Cdouble, Int32, Ptr{Ptr{Cvoid}}, Int32,
         Ptr{Cdouble}, Ptr{Ptr{Cfloat}}, Ptr{Cvoid}),
        h.ptr, gs, 0, nx, λv, Tv, λc, Tc, 1, recs, layout[], fill(Float64(GRID_TO_KELVIN), n), ptrs, C_NULL))
    check(ccall((:ceg_device_synchronize, libceg), Cint, (Int32,), device))
    out = EnergyGrid[]
    for g in 1:n
        iscoul = withcoul && g == n
        prec = iscoul ? ewald.precision : Inf
        host = Array{Cfloat,4}(undef, nz, ny, nx, 8)                                  # `.grid` units, for the file
        check(ccall((:ceg_device_download, libceg), Cint, (Int32, Ptr{Cvoid}, Int64, Ptr{Cvoid}), device, d[g][], bytes, host))
        check(ccall((:ceg_device_free, libceg), Cint, (Ptr{Cvoid},), d[g][]))
        eg = EnergyGrid(cset, num_unitcell, prec, true, Float32.(host .* GRID_TO_KELVIN))   # parse_grid's view (grids.jl:78)
        lock(GRID_LOCK) do
            GRID_HANDLES[eg.grid] = grids[g][]                                          # resident before any file exists
        end
        path = iscoul ? coulomb_path : paths[g]
        Threads.@spawn begin                                                            # the file, in the reference's byte layout
            mkpath(dirname(path))
            open(path, "w") do f
                _create_grid_common(f, cset, num_unitcell)
                iscoul && write(f, ewald.precision)
                write(f, host)
                write(f, NoUnits.(cset.cell.mat ./ u"Å"))
            end
        end
        RESIDENT[path] = eg
        push!(out, eg)
    end
    out
end

# ---- one process per GPU: construction and hand-off in one kernel (DESIGN §5b) ---------------------------------------------
# Every rank allocates its replica of the grids, the record buffers are exchanged once as CUDA IPC handles
# (`allgather(x)::Vector` = all-gather of a serialisable value over the ranks, indexed by rank + 1, e.g. MPI.jl's
# `MPI.Allgather` on the serialised bytes), and each rank's build kernel then writes its x-slab as finished
# interpolation records into the grids of ALL ranks over NVLink (ceg_grid_build_records).  After `build!` on every
# rank + a device synchronisation + a barrier, `grids` (one EnergyGrid per probe, Coulomb last) are complete on every
# GPU; the reference exchanges `.grid` files instead (raspa.jl:430-438).
mutable struct ShardedGridBuild
    probe::ProbeHandle
    cset::GridCoordinatesSetup
    xrange::UnitRange{Int}                 # this rank's planes (0-based, inclusive)
    handles::Vector{Ptr{Cvoid}}            # this rank's ceg_grid_h, one per grid
    dest::Vector{Ptr{Cvoid}}               # world x n_grids record pointers (rank-major)
    peers::Vector{Ptr{Cvoid}}              # ceg_ipc_open mappings, to close before the owners free their grids
    layout::Int32
    device::Int
end

function ShardedGridBuild(framework, forcefield::ForceField, spacing::TÅ, atoms::Vector{Symbol},
                          ewald::Union{Nothing,EwaldFramework}, rank::Int, world::Int, allgather; device=0)
    cset, _ = _setup_grid_common(framework, spacing, forcefield.cutoff)
    probe = ProbeSystem(framework, forcefield)
    idx = [forcefield.sdict[Symbol(CrystalEnergyGrids.get_atom_name(a))] for a in atoms]
    withcoul = !(ewald isa Nothing)
    n = length(idx) + withcoul
    h = ProbeHandle(probe, idx, withcoul, withcoul ? ustrip(u"Å^-1", ewald.α) : 0.0; device)
    cell = CegCell(SMatrix{3,3,Float64,9}(ustrip.(u"Å", cset.cell.mat)), SMatrix{3,3,Float64,9}(ustrip.(u"Å^-1", cset.cell.invmat)))
    gs = CegGridSpec(cset)
    handles = Ptr{Cvoid}[]; mine = Tuple{Vector{UInt8},Int64}[]; recs = Ptr{Cvoid}[]
    layout = Ref{Int32}(0)
    for g in 1:n
        gh = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:ceg_grid_create_empty, libceg), Cint, (Int32, Ref{CegCell}, Ref{CegGridSpec}, Int32, Ref{Ptr{Cvoid}}),
            device, cell, gs, !(withcoul && g == n), gh))
        r = Ref{Ptr{Cvoid}}(C_NULL)
        check(ccall((:ceg_grid_records, libceg), Cint, (Ptr{Cvoid}, Ref{Ptr{Cvoid}}, Ptr{Int64}, Ref{Int32}), gh[], r, C_NULL, layout))
        hd = Vector{UInt8}(undef, 64); off = Ref{Int64}(0)
        check(ccall((:ceg_ipc_export, libceg), Cint, (Int32, Ptr{Cvoid}, Ptr{UInt8}, Ref{Int64}), device, r[], hd, off))
        push!(handles, gh[]); push!(recs, r[]); push!(mine, (hd, off[]))
    end
    dest = Ptr{Cvoid}[]; peers = Ptr{Cvoid}[]
    for (r, theirs) in enumerate(world > 1 ? allgather(mine) : [mine])
        if r == rank + 1
            append!(dest, recs)
        else
            for (hd, off) in theirs
                p = Ref{Ptr{Cvoid}}(C_NULL)
                check(ccall((:ceg_ipc_open, libceg), Cint, (Int32, Ptr{UInt8}, Int64, Ref{Ptr{Cvoid}}), device, hd, off, p))
                push!(dest, p[]); push!(peers, p[])
            end
        end
    end
    nx = Int(cset.dims[1]) + 1
    lo = (nx * rank) ÷ world; hi = (nx * (rank + 1)) ÷ world - 1
    ShardedGridBuild(h, cset, lo:hi, handles, dest, peers, layout[], device)
end

# launch this rank's slab (asynchronous on the default stream); synchronise the device and the ranks before using the grids
function build!(sb::ShardedGridBuild)
    isempty(sb.xrange) && return sb
    n = length(sb.handles)
    λv = inv(GRID_TO_KELVIN); Tv = GRID_TO_KELVIN * 1e7
    λc = ustrip(u"K*Å/e_au^2", COULOMBIC_CONVERSION_FACTOR) / GRID_TO_KELVIN; Tc = inv(λc) * 1e7
    check(ccall((:ceg_grid_build_records, libceg), Cint,
        (Ptr{Cvoid}, Ref{CegGridSpec}, Int32, Int32, Cdouble, Cdouble, Cdouble, Cdouble, Int32, Ptr{Ptr{Cvoid}}, Int32,
         Ptr{Cdouble}, Ptr{Ptr{Cfloat}}, Ptr{Cvoid}),
        sb.probe.ptr, CegGridSpec(sb.cset), first(sb.xrange), last(sb.xrange) + 1, λv, Tv, λc, Tc,
        length(sb.dest) ÷ n, sb.dest, sb.layout, fill(Float64(GRID_TO_KELVIN), n), C_NULL, C_NULL))
    check(ccall((:ceg_device_synchronize, libceg), Cint, (Int32,), sb.device))
    sb
end

# unmap the peers' buffers (call on every rank, then barrier, before any rank frees its grids)
function close_peers!(sb::ShardedGridBuild)
    foreach(p -> check(ccall((:ceg_ipc_close, libceg), Cint, (Int32, Ptr{Cvoid}), sb.device, p)), sb.peers)
    empty!(sb.peers)
    sb
end

end # module
