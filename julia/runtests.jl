# julia/runtests.jl — the parity tests a maintainer runs on a machine with Julia AND a B200: the CUDA engine (julia/TNQSCuda.jl over
# libtnqs_cuda.so) against the REAL reference CPU path on identical inputs, at BASELINE.json's tolerances (truncation ranks exact,
# singular values / BP messages 1e-10 relative, <Z>, <ZZ> 1e-8 absolute).  This is what pins the engine to the reference itself
# (tests/ in this repository can only pin it to the NumPy restatement: "parity unpinned", DESIGN.md section 3).
# NOT EXECUTED in this repository's build image or on its GPU box (no Julia there).  Modelled on the reference's own tests
# (test/test_beliefpropagation.jl:26-53, test/test_apply.jl:11-53, test/test_expect.jl, test/test_truncate.jl:16-33).
#     julia --project=/path/to/reference julia/runtests.jl
using Test, Random, LinearAlgebra
using TensorNetworkQuantumSimulator, ITensors, NamedGraphs
const TN = TensorNetworkQuantumSimulator
include(joinpath(@__DIR__, "TNQSCuda.jl"))
using .TNQSCuda

relerr(a, b) = norm(a - b) / max(norm(b), 1.0e-300)

function message_array(bpc, e)                      # (b, b') array of the reference's message on the directed edge e
    b = only(TN.virtualinds(bpc, e))
    return ComplexF64.(ITensors.array(TN.message(bpc, e), b, b'))
end

function ising_layer(g; dt = 0.25, hx = 1.0, hz = 0.8, J = 0.5)
    layer = []
    append!(layer, ("Rx", [v], 2 * hx * dt) for v in vertices(g))
    append!(layer, ("Rz", [v], 2 * hz * dt) for v in vertices(g))
    for colored_edges in edge_color(g, 4)
        append!(layer, ("Rzz", pair, 2 * J * dt) for pair in colored_edges)
    end
    return layer
end

@testset "BP sweep: messages after fixed sweeps (reference order)" begin
    Random.seed!(123)
    for (g, χ) in ((named_comb_tree((3, 3)), 2), (named_grid((4, 4)), 3), (named_grid((4, 4)), 8))
        ψ = random_tensornetworkstate(ComplexF64, g; bond_dimension = χ)
        seq = TN.forest_cover_edge_sequence(g)
        ref = update(BeliefPropagationCache(ψ); maxiter = 3, tolerance = nothing, edge_sequence = seq)
        c = TNQSCuda.update!(CudaBeliefPropagationCache(ψ); maxiter = 3, tolerance = nothing, edge_sequence = seq)
        for e in seq
            b = only(TN.virtualinds(ref, e))
            got = ComplexF64.(ITensors.array(TN.message(c, e), b, b'))      # same Index objects: the cache keeps ψ's bond indices
            @test relerr(got, message_array(ref, e)) < 1.0e-10
        end
        @test isapprox(TN.partitionfunction(c), TN.partitionfunction(ref); rtol = 1.0e-9)
        v = first(center(g))
        @test abs(expect(c, ("Z", [v])) - expect(ref, ("Z", [v]))) < 1.0e-8
    end
end

@testset "BP on a tree is exact" begin                          # test/test_beliefpropagation.jl:47-53
    Random.seed!(123)
    g = named_comb_tree((3, 3))
    ψ = random_tensornetworkstate(ComplexF64, g; bond_dimension = 2)
    c = update(CudaBeliefPropagationCache(ψ))
    @test TN.partitionfunction(c) ≈ norm_sqr(ψ; alg = "exact")
end

@testset "apply_gates: ranks exact, two-site tensors, errors, <Z>, <ZZ>" begin        # test/test_apply.jl:22-53
    Random.seed!(123)
    g = named_grid((3, 3))
    ψ0 = normalize(random_tensornetworkstate(ComplexF64, g; bond_dimension = 1); alg = "bp")
    layer = ising_layer(g)
    for apply_kwargs in ((; cutoff = 1.0e-10, normalize_tensors = false), (; maxdim = 4, cutoff = 1.0e-10, normalize_tensors = true))
        ref, errs_ref = apply_gates(repeat(layer, 3), BeliefPropagationCache(ψ0) |> update; apply_kwargs)
        c, errs = TNQSCuda.apply_gates!(repeat(layer, 3), update(CudaBeliefPropagationCache(ψ0; chi_cap = 16)); apply_kwargs)
        χref = [dim(only(TN.virtualinds(ref, e))) for e in edges(g)]
        @test collect(bond_dims(c)) == χref                                                        # truncation ranks bit-exact
        @test isapprox(errs, errs_ref; rtol = 1.0e-8, atol = 1.0e-14)
        for v in vertices(g)
            @test abs(expect(c, ("Z", [v])) - expect(ref, ("Z", [v]))) < 1.0e-8
        end
        e = first(edges(g))
        @test abs(expect(c, ("ZZ", [src(e), dst(e)])) - expect(ref, ("ZZ", [src(e), dst(e)]))) < 1.0e-8
        # singular values of the last gates = the diagonal messages both sides leave on the edge (apply_gates.jl:108-114)
        ψ = network(c)
        @test maxvirtualdim(ψ) == maximum(χref)
    end
end

@testset "two-qubit known answer" begin                             # test/test_apply.jl:11-20
    circuit = [("Rx", [(1, 1)], 0.5), ("Rx", [(2, 1)], 0.2), ("CPHASE", [(1, 1), (2, 1)], -0.3)]
    g = build_graph_from_circuit(circuit)
    ψ0 = tensornetworkstate(ComplexF64, v -> "↓", g)
    ψ, _ = apply_gates(circuit, ψ0, Val(:cuda); apply_kwargs = (; maxdim = 2, cutoff = 1.0e-10, normalize_tensors = false))
    @test maxvirtualdim(ψ) <= 2
    @test norm_sqr(ψ; alg = "exact") ≈ 1.0
end

@testset "truncate, symmetric gauge, entanglement" begin            # test/test_truncate.jl:16-33, test/test_constructors.jl:69-73
    Random.seed!(123)
    g = named_grid((3, 3))
    ψ = random_tensornetworkstate(ComplexF64, g; bond_dimension = 4)
    ref = truncate(update(BeliefPropagationCache(ψ)); maxdim = 2)
    c = truncate(update(CudaBeliefPropagationCache(ψ)); maxdim = 2)
    @test maxvirtualdim(c) == 2
    for v in vertices(g)
        @test abs(expect(c, ("Z", [v])) - expect(ref, ("Z", [v]))) < 1.0e-8
    end
    cg = symmetric_gauge(update(CudaBeliefPropagationCache(ψ)))
    rg = symmetric_gauge(update(BeliefPropagationCache(ψ)))
    e = first(edges(g))
    @test abs(entanglement(cg, e) - entanglement(rg, e)) < 1.0e-8
end

@testset "inner(ψ, ϕ; alg = bp) and the Heisenberg picture" begin   # examples/2dIsing_dynamics_Heisenbergpicture.jl
    Random.seed!(123)
    g = named_comb_tree((3, 3))
    ψ = random_tensornetworkstate(ComplexF64, g; bond_dimension = 3)
    ϕ = random_tensornetworkstate(ComplexF64, g, siteinds(ψ); bond_dimension = 2)
    @test isapprox(inner(ψ, ϕ, Val(:cuda)), inner(ψ, ϕ; alg = "exact"); rtol = 1.0e-10)
    g = named_grid((4, 4))
    vz = first(center(g))
    O0 = paulitensornetworkstate(ComplexF64, v -> v == vz ? "Z" : "I", g)
    layer = []
    append!(layer, ("Rz", [v], -0.04) for v in vertices(g))
    for colored_edges in edge_color(g, 4)
        append!(layer, ("Rxx", pair, -0.08) for pair in colored_edges)
    end
    apply_kwargs = (; maxdim = 4, cutoff = 1.0e-14, normalize_tensors = false)
    ref, _ = apply_gates(layer, update(BeliefPropagationCache(O0)); apply_kwargs)
    c, _ = TNQSCuda.apply_gates!(layer, update(CudaBeliefPropagationCache(O0; chi_cap = 4)); apply_kwargs)
    @test collect(bond_dims(c)) == [dim(only(TN.virtualinds(ref, e))) for e in edges(g)]
    @test isapprox(inner(network(c), O0, Val(:cuda)), inner(network(ref), O0; alg = "bp"); atol = 1.0e-8)
end
