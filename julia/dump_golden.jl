# julia/dump_golden.jl — produces TRUE golden vectors from the real reference (TensorNetworkQuantumSimulator.jl) for the inputs
# of tests/golden/make_golden.py.  NOT RUN in this repository's build environment (Julia is not installed there); shipped so that
# anyone with Julia >= 1.10 and the reference installed can pin the oracle:  julia --project=/path/to/reference julia/dump_golden.jl
#
# Inputs are read from tests/golden/*_inputs.bin (raw ComplexF64, column-major, written by `python tests/golden/make_golden.py --inputs`)
# so that Julia's RNG is not involved; outputs are raw ComplexF64 / Float64 / Int32 arrays next to them.
using TensorNetworkQuantumSimulator, ITensors, NamedGraphs
const TN = TensorNetworkQuantumSimulator

function load_state(path, g, chi)
    s = siteinds("S=1/2", g)
    io = open(path); tensors = Dictionary{vertextype(g), ITensor}()
    es = collect(edges(g)); l = Dict(e => Index(chi) for e in es)
    for v in collect(vertices(g))
        legs = [l[e] for e in es if src(e) == v || dst(e) == v]       # ascending edge id = the ABI's canonical order
        n = 2 * chi^length(legs)
        A = reshape(reinterpret(ComplexF64, read(io, 16n)), (2, fill(chi, length(legs))...))
        set!(tensors, v, itensor(A, only(s[v]), legs...))
    end
    close(io)
    return TensorNetworkState(TensorNetwork(tensors, g), s)
end

function main()
    g = named_grid((3, 3))
    ψ = load_state(joinpath(@__DIR__, "..", "tests", "golden", "bp_grid3x3_chi3_inputs.bin"), g, 3)
    seq = TN.forest_cover_edge_sequence(g)
    bpc = update(BeliefPropagationCache(ψ); maxiter = 3, edge_sequence = seq)
    open(joinpath(@__DIR__, "..", "tests", "golden", "bp_grid3x3_chi3_julia.bin"), "w") do io
        for e in seq                                                    # messages in the reference's own edge order
            m = TN.message(bpc, e); b = only(TN.virtualinds(bpc, e))
            write(io, ComplexF64.(ITensors.array(m, b, b')))
        end
        for v in vertices(g); write(io, ComplexF64(expect(bpc, ("Z", [v])))); end
    end
    println("edge sequence (src => dst): ", seq)
end
main()
