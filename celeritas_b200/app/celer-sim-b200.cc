//---------------------------------------------------------------------------//
// celer-sim-b200: command-line front end with the reference app's calling convention
// (/root/reference/app/celer-sim/celer-sim.cc:146-260):
//     celer-sim-b200 {input}.json     run the input, print the JSON report to stdout
//     celer-sim-b200 -                read the input from stdin
// Links only against the C-ABI (include/celeritas_b200.h).
//---------------------------------------------------------------------------//
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>

#include "celeritas_b200.h"

static void print_usage(char const* exec_name)
{
    std::cerr << "usage: " << exec_name << " {input}.json\n"
              << "       " << exec_name << " [--help|-h]\n"
              << "       " << exec_name << " --version\n";
}

int main(int argc, char* argv[])
{
    if (argc != 2)
    {
        print_usage(argv[0]);
        return EXIT_FAILURE;
    }
    std::string filename{argv[1]};
    if (filename == "--help" || filename == "-h")
    {
        print_usage(argv[0]);
        return EXIT_SUCCESS;
    }
    if (filename == "--version" || filename == "-v")
    {
        std::cout << "celeritas_b200 0.1 (sm_100a)" << std::endl;
        return EXIT_SUCCESS;
    }
    std::stringstream text;
    if (filename == "-")
    {
        text << std::cin.rdbuf();
        filename = "<stdin>";
    }
    else
    {
        std::ifstream infile(filename);
        if (!infile)
        {
            std::cerr << "critical: Failed to open '" << filename << "'" << std::endl;
            return EXIT_FAILURE;
        }
        text << infile.rdbuf();
    }
    char* report = nullptr;
    int rc = b200_celer_sim_run(text.str().c_str(), &report);
    if (rc != B200_OK)
    {
        std::cerr << "critical: While running input at " << filename << ": " << b200_last_error()
                  << " (code " << rc << ")" << std::endl;
        // Same shape as the reference's ExceptionOutput: a report with the error
        std::cout << "{\"result\": {\"exception\": {\"code\": " << rc << "}}}" << std::endl;
        return EXIT_FAILURE;
    }
    std::cout << report << std::endl;
    b200_string_free(report);
    return EXIT_SUCCESS;
}
