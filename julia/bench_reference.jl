# julia/bench_reference.jl — times the REAL reference (TensorNetworkQuantumSimulator.jl, ITensors CPU path) on bench.py's workload
# (SURVEY.md section 8d, config C2: 32x32 open grid, chi = 16, ComplexF64; one Trotter layer = Rx, Rz on all sites, 4 Rzz colour
# classes with maxdim = 16, cutoff = 1e-10, lazy BP refreshes, <Z> at the centre) and prints the same JSON keys as
# `bench.py --impl reference`.  NOT RUN in this repository's build image or on its GPU box (no Julia there): bench.py times the
# NumPy restatement instead and labels it "kind": "port".  For anyone who has Julia:
#     julia --project=/path/to/reference -t auto julia/bench_reference.jl [nx ny chi steps]
# Random values come from Julia's RNG (throughput does not depend on them); parity inputs are the job of julia/dump_golden.jl.
using TensorNetworkQuantumSimulator, ITensors, NamedGraphs, LinearAlgebra, Random
const TN = TensorNetworkQuantumSimulator

function ising_layer(g; dt = 0.25, hx = 1.0, hz = 0.8, J = 0.5)          # examples/2dIsing_dynamics.jl:13-41
    layer = []
    append!(layer, ("Rx", [v], 2 * hx * dt) for v in vertices(g))
    append!(layer, ("Rz", [v], 2 * hz * dt) for v in vertices(g))
    for colored_edges in edge_color(g, 4)
        append!(layer, ("Rzz", pair, 2 * J * dt) for pair in colored_edges)
    end
    return layer
end

function main()
    nx, ny, χ, steps = (length(ARGS) >= 4 ? parse.(Int, ARGS[1:4]) : (32, 32, 16, 1))
    Random.seed!(1)
    g = named_grid((nx, ny))
    ψ = random_tensornetworkstate(ComplexF64, g; bond_dimension = χ)
    bpc = rescale(update(BeliefPropagationCache(ψ); maxiter = 25, tolerance = 1.0e-8))
    layer = ising_layer(g)
    apply_kwargs = (; maxdim = χ, cutoff = 1.0e-10, normalize_tensors = true)
    nmsg = 2 * length(edges(g))

    # BP edge-message updates/s: one reference-order sweep (update_iteration!, abstractbeliefpropagationcache.jl:182-196)
    seq = TN.forest_cover_edge_sequence(g)
    update(bpc; maxiter = 1, tolerance = nothing, edge_sequence = seq)      # warm-up (compilation)
    t_bp = @elapsed update(bpc; maxiter = 1, tolerance = nothing, edge_sequence = seq)

    # whole layers through the public API (apply_gates + expect), as bench.py's e2e leg
    apply_gates(layer, bpc; apply_kwargs)                                   # warm-up
    t_layer = @elapsed for _ in 1:steps
        bpc, errs = apply_gates(layer, bpc; apply_kwargs)
        expect(bpc, ("Z", [(nx ÷ 2, ny ÷ 2)]))
    end
    q(x) = "\"" * string(x) * "\""                                             # JSON string
    fields = [q("impl") * ": " * q("reference"), q("metric") * ": " * q("bp_edge_message_updates_per_sec"),
              q("value") * ": " * string(nmsg / t_bp), q("unit") * ": " * q("messages/s"),
              q("ms_per_step") * ": " * string(1000 * t_layer / steps),
              q("cpu_baseline") * ": {" * q("kind") * ": " * q("reference") * ", " * q("cores") * ": " * string(Threads.nthreads()) * ", " *
                  q("blas_threads") * ": " * string(BLAS.get_num_threads()) * ", " *
                  q("sample") * ": " * q("one reference-order sweep of $(nmsg) messages; $(steps) full layer(s)") * "}",
              q("config") * ": {" * q("workload") * ": " * q("C2 $(nx)x$(ny) grid chi=$(χ) ComplexF64") * "}"]
    println("{" * join(fields, ", ") * "}")
end
main()
