// Times the reference's sliced-benchmark flow through the drop-in headers: load a network file, make one
// SliceIndices copy per slice, AddContractionTasks for each, AddReductionTask, Contract().wait() — the call
// sequence of /root/reference/examples/paper_benchmarks/CPU/jet_cpu_m10/jet_sliced.cpp:33-98 (this file is a
// re-statement with a JSON report, repetitions and a slice-count cap, not that source).  `api` = "tbc" runs the
// flow above, "sliced" the SlicedContractor on the same slices for comparison.
//
//   tbc_bench <file.json> <comma-separated sliced indices> [--api tbc|sliced] [--slices N] [--reps R] [--c128]
//
// Prints one JSON line: seconds of Contract().wait() (best and all repetitions), host preparation time, result.
#include <chrono>
#include <complex>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>

#include "Jet.hpp"

using namespace Jet;

template <class T> int Run(const std::string &file_name, const std::vector<std::string> &sliced, const std::string &api,
                           size_t max_slices, int reps)
{
    using tensor_t = Tensor<T>;
    using clock = std::chrono::steady_clock;
    std::ifstream in(file_name);
    if (!in) {
        std::cerr << "cannot open " << file_name << std::endl;
        return 2;
    }
    std::string text{std::istreambuf_iterator<char>(in), std::istreambuf_iterator<char>()};
    const auto file = TensorNetworkSerializer<tensor_t>()(text);
    const auto path = file.path.value().GetPath();
    size_t num_slices = 1;
    for (const auto &index : sliced)
        num_slices *= file.tensors.GetIndexToEdgeMap().at(index).dim;
    const size_t count = std::min(num_slices, max_slices);

    std::vector<double> seconds;
    double prep = 0;
    std::complex<double> result{0, 0};
    size_t shared = 0;
    for (int r = 0; r < reps; r++) {
        if (api == "tbc") {
            const auto p0 = clock::now();
            std::vector<TensorNetwork<tensor_t>> slices(count);
            for (size_t i = 0; i < count; i++) {
                slices[i] = file.tensors;
                slices[i].SliceIndices(sliced, i);
            }
            TaskBasedContractor<tensor_t> contractor;
            shared = 0;
            for (size_t i = 0; i < count; i++) {
                PathInfo pinfo(slices[i], path);
                shared += contractor.AddContractionTasks(slices[i], pinfo);
            }
            contractor.AddReductionTask();
            const auto t1 = clock::now();
            contractor.Contract().get();
            const auto t2 = clock::now();
            prep = std::chrono::duration<double>(t1 - p0).count();
            seconds.push_back(std::chrono::duration<double>(t2 - t1).count());
            result = contractor.GetReductionResult().GetSize() == 1 ? std::complex<double>(contractor.GetReductionResult().GetScalar())
                                                                   : std::complex<double>(contractor.GetReductionResult().GetData()[0]);
        }
        else {
            const auto t1 = clock::now();
            SlicedContractor<tensor_t> sc(file.tensors, path, sliced, 0, 0, 0);
            const auto out = sc.Contract(0, count);
            const auto t2 = clock::now();
            seconds.push_back(std::chrono::duration<double>(t2 - t1).count());
            result = out.GetSize() == 1 ? std::complex<double>(out.GetScalar()) : std::complex<double>(out.GetData()[0]);
        }
    }
    double best = seconds[0];
    for (double s : seconds)
        best = std::min(best, s);
    std::ostringstream os;
    os.precision(17);
    os << "{\"api\": \"" << api << "\", \"file\": \"" << file_name << "\", \"dtype\": \"" << (sizeof(T) == 8 ? "c64" : "c128")
       << "\", \"num_sliced\": " << sliced.size() << ", \"slices\": " << count << ", \"shared_tasks\": " << shared
       << ", \"prep_s\": " << prep << ", \"contract_s\": " << best << ", \"slices_per_s\": " << count / best << ", \"all_s\": [";
    for (size_t i = 0; i < seconds.size(); i++)
        os << (i ? ", " : "") << seconds[i];
    os << "], \"result\": [" << result.real() << ", " << result.imag() << "]}";
    std::cout << os.str() << std::endl;
    return 0;
}

int main(int argc, char **argv)
{
    if (argc < 3) {
        std::cerr << "usage: tbc_bench <file.json> <idx,idx,...> [--api tbc|sliced] [--slices N] [--reps R] [--c128]" << std::endl;
        return 2;
    }
    std::vector<std::string> sliced;
    {
        std::stringstream ss(argv[2]);
        std::string item;
        while (std::getline(ss, item, ','))
            if (!item.empty())
                sliced.push_back(item);
    }
    std::string api = "tbc";
    size_t max_slices = static_cast<size_t>(-1);
    int reps = 3;
    bool c128 = false;
    for (int i = 3; i < argc; i++) {
        if (!std::strcmp(argv[i], "--api") && i + 1 < argc)
            api = argv[++i];
        else if (!std::strcmp(argv[i], "--slices") && i + 1 < argc)
            max_slices = std::stoull(argv[++i]);
        else if (!std::strcmp(argv[i], "--reps") && i + 1 < argc)
            reps = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--c128"))
            c128 = true;
    }
    try {
        return c128 ? Run<std::complex<double>>(argv[1], sliced, api, max_slices, reps)
                    : Run<std::complex<float>>(argv[1], sliced, api, max_slices, reps);
    }
    catch (const std::exception &e) {
        std::cerr << "error: " << e.what() << std::endl;
        return 1;
    }
}
