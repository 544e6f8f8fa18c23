// Minimal stand-in for the reference's Config (src/config.h:53-110, src/config.cpp) for builds
// outside the reference tree: INI file + `--section:key=value` command-line overrides
// (src/config.cpp:122-150) and defaulted typed getters (src/config.cpp:255-297).
// Inside the reference tree the real "config.h" is used instead.
#pragma once
#include <cstdio>
#include <cstring>
#include <map>
#include <sstream>
#include <string>

class Config {
    std::map<std::string, std::map<std::string, std::string>> data;

    static std::string trim(const std::string& s)
    {
        size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
        return a == std::string::npos ? "" : s.substr(a, b - a + 1);
    }
    template <typename V> V lookup(const std::string& sec, const std::string& name, V def) const
    {
        auto s = data.find(sec);
        if (s == data.end()) return def;
        auto k = s->second.find(name);
        if (k == s->second.end()) return def;
        std::istringstream is(k->second);
        V v = def;
        is >> v;
        return is.fail() ? def : v;
    }

public:
    void set(const std::string& sec, const std::string& name, const std::string& value) { data[sec][name] = value; }

    // a missing file is silently fine, like the reference (src/config.cpp:68-71)
    void open(const std::string& filename)
    {
        FILE* f = fopen(filename.c_str(), "r");
        if (!f) return;
        char line[4096];
        std::string sec;
        while (fgets(line, sizeof(line), f)) {
            std::string s = trim(line);
            if (s.empty() || s[0] == ';' || s[0] == '#') continue;
            if (s.front() == '[' && s.back() == ']') { sec = trim(s.substr(1, s.size() - 2)); continue; }
            size_t eq = s.find('=');
            if (eq == std::string::npos) continue;
            data[sec][trim(s.substr(0, eq))] = trim(s.substr(eq + 1));
        }
        fclose(f);
    }

    void rewrite(int argc, char** argv)
    {
        for (int i = 1; i < argc; i++) {
            const char* a = argv[i];
            if (strncmp(a, "--", 2) != 0) continue;
            std::string s(a + 2);
            size_t c = s.find(':'), eq = s.find('=');
            if (c == std::string::npos || eq == std::string::npos || eq < c) continue;
            data[s.substr(0, c)][s.substr(c + 1, eq - c - 1)] = s.substr(eq + 1);
        }
    }

    double get(const std::string& sec, const std::string& name, double def) const { return lookup<double>(sec, name, def); }
    int get(const std::string& sec, const std::string& name, int def) const { return lookup<int>(sec, name, def); }
    std::string get(const std::string& sec, const std::string& name, const std::string& def) const
    {
        auto s = data.find(sec);
        if (s == data.end()) return def;
        auto k = s->second.find(name);
        return k == s->second.end() ? def : k->second;
    }
    std::string get(const std::string& sec, const std::string& name, const char* def) const { return get(sec, name, std::string(def)); }
};
