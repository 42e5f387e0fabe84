// smallk_b200 host — result files of flat clustering, same names and text as the reference:
//   assignments_flat_N.csv (common/src/assignments.cpp:21-40), assignments_fuzzy_N.csv (:43-70, 3 digits),
//   clusters_N.{xml,json} (common/src/flat_clust_output.cpp:49-173, flatclust_xml_writer.cpp, flatclust_json_writer.cpp),
// and the dictionary loader (one term per line).
#pragma once

#include <fstream>
#include <iomanip>
#include <iostream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

#include "hierclust_writer.hpp"

inline std::string EnsureTrailingPathSep(const std::string& dir)
{
    if (dir.empty() || dir.back() == '/') return dir;
    return dir + "/";
}

inline bool LoadStringsFromFile(const std::string& path, std::vector<std::string>& out)
{
    // common/src/utils.cpp:220-239, rule for rule: lines are APPENDED to `out`, a carriage return stays part of the string,
    // blank lines are strings, and a last line that is not terminated by a newline is not read
    // (tests/test_file_formats.py::test_dictionary_reader_matches_reference).
    std::ifstream in(path);
    if (!in) return false;
    std::string line;
    while (std::getline(in, line) && !in.eof()) out.push_back(line);
    return true;
}

inline bool WriteAssignmentsFile(const std::vector<unsigned int>& labels, const std::string& path)
{
    std::ofstream out(path);
    if (!out) return false;
    for (size_t i = 0; i < labels.size(); ++i) { if (i) out << ','; out << labels[i]; }
    out << std::endl;
    return true;
}

inline bool WriteFuzzyAssignmentsFile(const std::vector<float>& prob, const unsigned int k, const unsigned int n, const std::string& path)
{
    std::ofstream out(path);
    if (!out) return false;
    out << std::scientific << std::setprecision(3);
    for (unsigned int c = 0; c < n; ++c)
    {
        for (unsigned int r = 0; r < k; ++r) { if (r) out << ','; out << prob[static_cast<size_t>(c) * k + r]; }
        out << std::endl;
    }
    return true;
}

inline void FlatClustWriteResults(const std::string& assignfile, const std::string& fuzzyfile, const std::string& resultfile,
                                  const std::vector<unsigned int>& assignments, const std::vector<float>& probabilities,
                                  const std::vector<std::string>& dictionary, const std::vector<int>& term_indices,
                                  const FileFormat format, const unsigned int maxterms, const unsigned int num_docs,
                                  const unsigned int num_clusters)
{
    if (term_indices.size() < static_cast<size_t>(num_clusters) * maxterms) throw std::logic_error("FlatClustWriteResults: term count is invalid");
    if (FileFormat::XML != format && FileFormat::JSON != format) throw std::logic_error("CreateFlatclustWriter: invalid file format");
    std::map<int, int> doc_counts;
    for (unsigned int a : assignments) doc_counts[static_cast<int>(a)] += 1;
    if (doc_counts.size() != num_clusters)
        std::cout << "Warning: only " << doc_counts.size() << " clusters received an assignment." << std::endl << std::endl;
    if (!WriteAssignmentsFile(assignments, assignfile)) std::cerr << "\terror writing flat assignments file" << std::endl;
    if (!WriteFuzzyAssignmentsFile(probabilities, num_clusters, num_docs, fuzzyfile)) std::cerr << "\terror writing fuzzy assignments file" << std::endl;
    std::ofstream out(resultfile);
    if (!out) { std::cerr << "FlatClustWriteResults: could not open output file " << resultfile << std::endl; return; }
    const bool xml = (FileFormat::XML == format);
    const std::string S4(4, ' '), S8(8, ' '), S12(12, ' '), S16(16, ' ');
    if (xml) out << "<?xml version=\"1.0\"?>" << std::endl << "<DataSet id=\"" << num_docs << "\">" << std::endl;
    else out << "{" << std::endl << S4 << "\"doc_count\": " << num_docs << "," << std::endl << S4 << "\"nodes\": [" << std::endl;
    for (unsigned int i = 0; i < num_clusters; ++i)
    {
        const auto it = doc_counts.find(static_cast<int>(i));
        const int count = (doc_counts.end() == it) ? 0 : it->second;
        if (xml)
        {
            out << S4 << "<node id=\"" << i << "\">" << std::endl << S8 << "<doc_count>" << count << "</doc_count>" << std::endl;
            if (count)
            {
                out << S8 << "<top_terms>" << std::endl;
                for (unsigned int q = 0; q < maxterms; ++q) out << S12 << "<term name=\"" << dictionary[term_indices[i * maxterms + q]] << "\"/>" << std::endl;
                out << S8 << "</top_terms>" << std::endl;
            }
            out << S4 << "</node>" << std::endl;
        }
        else
        {
            if (i) out << "," << std::endl;
            out << S8 << "{" << std::endl << S12 << "\"id\": " << i << "," << std::endl << S12 << "\"doc_count\": " << count << "," << std::endl;
            if (count && maxterms)
            {
                out << S12 << "\"top_terms\": [" << std::endl;
                for (unsigned int q = 0; q + 1 < maxterms; ++q) out << S16 << "\"" << dictionary[term_indices[i * maxterms + q]] << "\"," << std::endl;
                out << S16 << "\"" << dictionary[term_indices[i * maxterms + maxterms - 1]] << "\"" << std::endl << S12 << "]" << std::endl;
            }
            out << S8 << "}";
        }
    }
    if (xml) out << "</DataSet>" << std::endl;
    else out << std::endl << S4 << "]" << std::endl << "}" << std::endl;
}

inline void FlatClustWriteResults(const std::string& outdir, const std::vector<unsigned int>& assignments,
                                  const std::vector<float>& probabilities, const std::vector<std::string>& dictionary,
                                  const std::vector<int>& term_indices, const FileFormat format, const unsigned int maxterms,
                                  const unsigned int num_docs, const unsigned int num_clusters)
{
    const std::string dir = EnsureTrailingPathSep(outdir);
    std::ostringstream a, f, r;
    a << "assignments_flat_" << num_clusters;
    f << "assignments_fuzzy_" << num_clusters;
    r << "clusters_" << num_clusters;
    FlatClustWriteResults(dir + AppendExtension(a.str(), FileFormat::CSV), dir + AppendExtension(f.str(), FileFormat::CSV),
                          dir + AppendExtension(r.str(), format), assignments, probabilities, dictionary, term_indices, format,
                          maxterms, num_docs, num_clusters);
}
