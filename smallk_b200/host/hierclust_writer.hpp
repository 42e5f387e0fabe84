// smallk_b200 host — tree-file writers for hierclust: the reference's IHierclustWriter interface
// (hierclust/include/hierclust_writer.hpp:21-44) with XML and JSON implementations that emit the same text as
// hierclust/src/hierclust_xml_writer.cpp:28-123 and hierclust_json_writer.cpp:28-146, and the factory of
// hierclust/src/hierclust_writer_factory.cpp.
#pragma once

#include <fstream>
#include <string>
#include <vector>

enum class FileFormat { CSV, XML, JSON, TXT };

inline std::string AppendExtension(const std::string& name, const FileFormat fmt)
{
    switch (fmt)
    {
    case FileFormat::CSV: return name + ".csv";
    case FileFormat::XML: return name + ".xml";
    case FileFormat::JSON: return name + ".json";
    default: return name + ".txt";
    }
}

class IHierclustWriter
{
public:
    virtual ~IHierclustWriter() {}
    virtual void WriteHeader(std::ofstream& outfile, const int doc_count) = 0;
    virtual void WriteNodeBegin(std::ofstream& outfile, const int node_id) = 0;
    virtual void WriteParentId(std::ofstream& outfile, const int parent_id) = 0;
    virtual void WriteLeftChild(std::ofstream& outfile, const bool is_left_child, const int lc_label) = 0;
    virtual void WriteRightChild(std::ofstream& outfile, const int rc_label) = 0;
    virtual void WriteDocCount(std::ofstream& outfile, const int count) = 0;
    virtual void WriteTopTerms(std::ofstream& outfile, const std::vector<int>& term_indices,
                               const std::vector<std::string>& dictionary) = 0;
    virtual void WriteNodeEnd(std::ofstream& outfile) = 0;
    virtual void WriteFooter(std::ofstream& outfile) = 0;
};

class HierclustXmlWriter : public IHierclustWriter
{
    static const char* pad(int levels) { static const char* p[] = {"", "    ", "        ", "            "}; return p[levels]; }
public:
    void WriteHeader(std::ofstream& o, const int doc_count) override
    { o << "<?xml version=\"1.0\"?>" << std::endl << "<DataSet id=\"" << doc_count << "\">" << std::endl; }
    void WriteNodeBegin(std::ofstream& o, const int id) override { o << pad(1) << "<node id=\"" << id << "\">" << std::endl; }
    void WriteParentId(std::ofstream& o, const int id) override { o << pad(2) << "<parent_id>" << id << "</parent_id>" << std::endl; }
    void WriteLeftChild(std::ofstream& o, const bool is_left, const int id) override
    {
        o << pad(2) << "<left_child>" << (is_left ? "true" : "false") << "</left_child>" << std::endl;
        o << pad(2) << "<left_child_id>" << id << "</left_child_id>" << std::endl;
    }
    void WriteRightChild(std::ofstream& o, const int id) override { o << pad(2) << "<right_child_id>" << id << "</right_child_id>" << std::endl; }
    void WriteDocCount(std::ofstream& o, const int count) override { o << pad(2) << "<doc_count>" << count << "</doc_count>" << std::endl; }
    void WriteTopTerms(std::ofstream& o, const std::vector<int>& terms, const std::vector<std::string>& dict) override
    {
        o << pad(2) << "<top_terms>" << std::endl;
        for (int t : terms) o << pad(3) << "<term name=\"" << dict[t] << "\"/>" << std::endl;
        o << pad(2) << "</top_terms>" << std::endl;
    }
    void WriteNodeEnd(std::ofstream& o) override { o << pad(1) << "</node>" << std::endl; }
    void WriteFooter(std::ofstream& o) override { o << "</DataSet>" << std::endl; }
};

class HierclustJsonWriter : public IHierclustWriter
{
    unsigned int written_ = 0;
    static std::string pad(int levels) { return std::string(4 * levels, ' '); }
public:
    void WriteHeader(std::ofstream& o, const int doc_count) override
    {
        o << "{" << std::endl << pad(1) << "\"doc_count\": " << doc_count << "," << std::endl << pad(1) << "\"nodes\": [" << std::endl;
        written_ = 0;
    }
    void WriteNodeBegin(std::ofstream& o, const int id) override
    {
        if (written_ > 0) o << "," << std::endl;
        o << pad(2) << "{" << std::endl << pad(3) << "\"id\": " << id << "," << std::endl;
    }
    void WriteParentId(std::ofstream& o, const int id) override { o << pad(3) << "\"parent_id\": " << id << "," << std::endl; }
    void WriteLeftChild(std::ofstream& o, const bool is_left, const int id) override
    {
        o << pad(3) << "\"left_child\": " << (is_left ? "true" : "false") << "," << std::endl;
        o << pad(3) << "\"left_child_id\": " << id << "," << std::endl;
    }
    void WriteRightChild(std::ofstream& o, const int id) override { o << pad(3) << "\"right_child_id\": " << id << "," << std::endl; }
    void WriteDocCount(std::ofstream& o, const int count) override { o << pad(3) << "\"doc_count\": " << count << "," << std::endl; }
    void WriteTopTerms(std::ofstream& o, const std::vector<int>& terms, const std::vector<std::string>& dict) override
    {
        if (terms.empty()) return;
        o << pad(3) << "\"top_terms\": [" << std::endl;
        for (size_t q = 0; q + 1 < terms.size(); ++q) o << pad(4) << "\"" << dict[terms[q]] << "\"," << std::endl;
        o << pad(4) << "\"" << dict[terms.back()] << "\"" << std::endl << pad(3) << "]" << std::endl;
    }
    void WriteNodeEnd(std::ofstream& o) override { o << pad(2) << "}"; ++written_; }
    void WriteFooter(std::ofstream& o) override { o << std::endl << pad(1) << "]" << std::endl << "}" << std::endl; }
};

inline IHierclustWriter* CreateHierclustWriter(const FileFormat& format)
{
    if (FileFormat::XML == format) return new HierclustXmlWriter();
    if (FileFormat::JSON == format) return new HierclustJsonWriter();
    return nullptr;
}
