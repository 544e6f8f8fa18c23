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
        // same line grammar as Config::load (src/config.cpp:78-120): "[section]" at the start of a line; lines that
        // start with '#' or ';' are comments; otherwise "key = value" where an inline ";" starts a comment and key and
        // value are the first two tokens separated by blanks, tabs or '='; lines before the first section are ignored
        char line[4096];
        std::string sec;
        bool have_sec = false;
        while (fgets(line, sizeof(line), f)) {
            std::string s(line);
            if (s[0] == '[') {
                size_t e = s.find_first_of(" \t\r\n", 1);
                std::string name = s.substr(1, (e == std::string::npos ? s.size() : e) - 1);
                if (!name.empty()) {                       // sscanf("[%s]") + drop the last character
                    name.pop_back();
                    sec = name; have_sec = true; data[sec];
                    continue;
                }
            }
            if (s[0] == '#' || s[0] == ';' || !have_sec) continue;
            size_t c = s.find(';');
            if (c != std::string::npos) s.resize(c);
            const char* sep = " =\t\r\n";
            size_t k0 = s.find_first_not_of(sep);
            if (k0 == std::string::npos) continue;
            size_t k1 = s.find_first_of(sep, k0);
            if (k1 == std::string::npos) continue;
            size_t v0 = s.find_first_not_of(sep, k1);
            if (v0 == std::string::npos) continue;
            size_t v1 = s.find_first_of(sep, v0);
            data[sec][s.substr(k0, k1 - k0)] = s.substr(v0, v1 == std::string::npos ? std::string::npos : v1 - v0);
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
