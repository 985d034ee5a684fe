// kami::options -- same surface and file syntax as the reference (kami/options.h:5-18,
// kami/options.cpp:89-143): a process-wide string map with typed getters, `key: value` lines,
// `#` comments.  Header-only here; the hot path reads its keys once in constructors.
#pragma once
#include <cctype>
#include <cerrno>
#include <cstring>
#include <fstream>
#include <iostream>
#include <map>
#include <mutex>
#include <stdexcept>
#include <string>

namespace kami::options {
namespace detail {
inline std::map<std::string, std::string>& values() {
    static std::map<std::string, std::string> v;
    return v;
}
inline std::mutex& lock() {
    static std::mutex m;
    return m;
}
}  // namespace detail

inline void setStr(std::string key, std::string value) {
    std::lock_guard<std::mutex> g(detail::lock());
    detail::values()[key] = value;
}
inline void setInt(std::string key, int value) { setStr(key, std::to_string(value)); }
inline void setFloat(std::string key, float value) { setStr(key, std::to_string(value)); }

inline std::string getStr(std::string key, std::string def = "") {
    std::lock_guard<std::mutex> g(detail::lock());
    auto it = detail::values().find(key);
    return it == detail::values().end() ? def : it->second;
}
inline int getInt(std::string key, int def = 0) {
    std::string s = getStr(key, std::to_string(def));
    try {
        return std::stoi(s);
    } catch (std::exception& e) {
        throw std::runtime_error("conversion failure for key \"" + key + "\" = \"" + s + "\": " + e.what());
    }
}
inline float getFloat(std::string key, float def = 0) {
    std::string s = getStr(key, std::to_string(def));
    try {
        return std::stof(s);
    } catch (std::exception& e) {
        throw std::runtime_error("conversion failure for key \"" + key + "\" = \"" + s + "\": " + e.what());
    }
}
inline void write(std::string path = "options.yml") {
    std::lock_guard<std::mutex> g(detail::lock());
    std::ofstream f(path);
    if (!f) throw std::runtime_error("couldn't open " + path + " for writing: " + strerror(errno));
    for (auto& kv : detail::values()) f << kv.first << ": " << kv.second << std::endl;
}
inline void print() {
    std::lock_guard<std::mutex> g(detail::lock());
    for (auto& kv : detail::values()) std::cout << kv.first << ": " << kv.second << std::endl;
}
inline void load(std::string path = "options.yml") {
    std::lock_guard<std::mutex> g(detail::lock());
    std::ifstream f(path);
    if (!f) throw std::runtime_error("couldn't open " + path + " for reading: " + strerror(errno));
    auto trim = [](std::string& s) {
        size_t a = 0, b = s.size();
        while (a < b && isspace((unsigned char)s[a])) ++a;
        while (b > a && isspace((unsigned char)s[b - 1])) --b;
        s = s.substr(a, b - a);
    };
    std::cout << "Importing options from " << path << std::endl;
    std::string line;
    int no = 0;
    while (std::getline(f, line)) {
        ++no;
        size_t h = line.find('#');
        if (h != std::string::npos) line.erase(h);
        size_t c = line.find(':');
        if (c == std::string::npos) continue;
        if (c == 0) throw std::runtime_error("invalid key at " + path + ":" + std::to_string(no));
        std::string key = line.substr(0, c), val = line.substr(c + 1);
        trim(key);
        trim(val);
        if (key.empty() || val.empty()) throw std::runtime_error("invalid option at " + path + ":" + std::to_string(no));
        std::cout << key << ": " << val << std::endl;
        detail::values()[key] = val;
    }
}
}  // namespace kami::options
