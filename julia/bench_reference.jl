# Times the REAL reference (atoptima/DynamicSparseArrays.jl) on bench.py's config-2 workload, read from the binary files
# written by tools/write_inputs.py.  Untested here (no Julia in the build image); it is the recipe SURVEY.md §8d asks for
# when a box has Julia:
#     julia -e 'using Pkg; Pkg.develop(path="/path/to/DynamicSparseArrays.jl")'
#     julia julia/bench_reference.jl OUTDIR [warmup]
# One step = a loop of setindex! over the batch (the reference has no batched update, matrix.jl:119-121) + A * x with x a
# full DynamicSparseVector (operations.jl:14-24).  Prints one JSON line in the format of `bench.py --impl reference`.
using DynamicSparseArrays

function readbin(T, path)
    n = filesize(path) ÷ sizeof(T)
    v = Vector{T}(undef, n)
    read!(path, v)
    return v
end

function main()
    dir = ARGS[1]
    warmup = length(ARGS) > 1 ? parse(Int, ARGS[2]) : 1
    m, n, nnz0, batch, nsteps = parse.(Int, split(readline(joinpath(dir, "meta.txt"))))
    I = readbin(Int64, joinpath(dir, "I.bin")); J = readbin(Int64, joinpath(dir, "J.bin")); V = readbin(Float64, joinpath(dir, "V.bin"))
    x = readbin(Float64, joinpath(dir, "x.bin"))
    A = dynamicsparse(I, J, V, m, n)                         # matrix.jl:15-19
    xs = dynamicsparsevec(collect(1:n), x)                   # every entry stored
    times = Float64[]
    y = nothing
    for s in 0:nsteps-1
        bi = readbin(Int64, joinpath(dir, "batch_$(lpad(s, 3, '0'))_i.bin"))
        bj = readbin(Int64, joinpath(dir, "batch_$(lpad(s, 3, '0'))_j.bin"))
        bv = readbin(Float64, joinpath(dir, "batch_$(lpad(s, 3, '0'))_v.bin"))
        t = @elapsed begin
            for k in eachindex(bv)
                A[bi[k], bj[k]] = bv[k]                      # matrix.jl:43-62
            end
            y = A * xs                                       # operations.jl:14-24
        end
        s >= warmup && push!(times, t)
    end
    total = sum(times)
    val = batch * length(times) / total / 1e6
    println("{\"impl\": \"reference\", \"metric\": \"batched PCSR insert/delete Mupdates/s\", \"value\": $val, \"unit\": \"Mupdates/s\", ",
            "\"steps\": $(length(times)), \"warmup\": $warmup, \"ms_per_step\": $(1e3 * total / length(times)), \"higher_is_better\": true, ",
            "\"cpu_baseline\": {\"value\": $val, \"unit\": \"Mupdates/s\", \"cores\": 1, \"kind\": \"reference\", ",
            "\"sample\": \"$(length(times)) full steps, Julia $(VERSION)\"}, \"checksum\": $(sum(values(y.nzval)))}")
end

main()
