// Small functional stand-in for the boost::program_options names /root/reference/main.cpp uses (TEST INFRASTRUCTURE):
// "--long value" / "-s value" options with defaults and required(), flags without value, variables_map::count / as<T>().
#ifndef APDE_STUB_BOOST_PO_HPP_
#define APDE_STUB_BOOST_PO_HPP_
#include <exception>
#include <map>
#include <memory>
#include <ostream>
#include <sstream>
#include <string>
#include <vector>
namespace boost { namespace program_options {
struct error : std::exception {
    std::string msg;
    explicit error(const std::string &m = "program_options error") : msg(m) {}
    const char *what() const noexcept override { return msg.c_str(); }
};
struct value_semantic {
    bool is_required = false, has_default = false;
    std::string default_text;
    virtual ~value_semantic() {}
};
template <class T> struct typed_value : value_semantic {
    typed_value *required() { is_required = true; return this; }
    typed_value *default_value(const T &v) {
        std::ostringstream os;
        os << std::boolalpha << v;
        default_text = os.str();
        has_default = true;
        return this;
    }
};
template <class T> typed_value<T> *value() { return new typed_value<T>(); }  // owned by the option table (leaked at exit, as a test stub may)
struct option_entry { std::string long_name, short_name, help; std::shared_ptr<value_semantic> sem; };
struct options_description;
struct options_description_easy_init {
    options_description *owner;
    inline options_description_easy_init &operator()(const char *name, value_semantic *v, const char *help);
    inline options_description_easy_init &operator()(const char *name, const char *help);
};
struct options_description {
    std::string caption;
    std::vector<option_entry> entries;
    explicit options_description(const char *c) : caption(c) {}
    options_description_easy_init add_options() { return options_description_easy_init{this}; }
    const option_entry *find(const std::string &key) const {
        for (auto &e : entries) if (("--" + e.long_name) == key || (!e.short_name.empty() && ("-" + e.short_name) == key)) return &e;
        return nullptr;
    }
};
inline void add_entry(options_description *d, const char *name, value_semantic *v, const char *help) {
    option_entry e;
    const std::string n(name);
    const size_t comma = n.find(',');
    e.long_name = n.substr(0, comma);
    if (comma != std::string::npos) e.short_name = n.substr(comma + 1);
    e.help = help;
    e.sem.reset(v);
    d->entries.push_back(e);
}
inline options_description_easy_init &options_description_easy_init::operator()(const char *name, value_semantic *v, const char *help) {
    add_entry(owner, name, v, help);
    return *this;
}
inline options_description_easy_init &options_description_easy_init::operator()(const char *name, const char *help) {
    add_entry(owner, name, nullptr, help);
    return *this;
}
inline std::ostream &operator<<(std::ostream &o, const options_description &d) {
    o << d.caption << ":\n";
    for (auto &e : d.entries) o << "  --" << e.long_name << (e.short_name.empty() ? "" : " [ -" + e.short_name + " ]") << "  " << e.help << "\n";
    return o;
}
struct variable_value {
    std::string text;
    template <class T> T as() const {
        T v{};
        std::istringstream is(text);
        is >> v;
        return v;
    }
};
template <> inline std::string variable_value::as<std::string>() const { return text; }
template <> inline bool variable_value::as<bool>() const {
    return text == "true" || text == "1" || text == "on" || text == "yes" || text == "True" || text == "TRUE";
}
struct variables_map {
    std::map<std::string, variable_value> values;
    const options_description *desc = nullptr;
    int count(const char *k) const { return values.count(k) ? 1 : 0; }
    const variable_value &operator[](const char *k) const {
        static const variable_value empty;
        auto it = values.find(k);
        return it == values.end() ? empty : it->second;
    }
};
struct parsed_options { std::map<std::string, std::string> given; const options_description *desc; };
inline parsed_options parse_command_line(int argc, char **argv, const options_description &d) {
    parsed_options p;
    p.desc = &d;
    for (int i = 1; i < argc; ++i) {
        const std::string key(argv[i]);
        const option_entry *e = d.find(key);
        if (!e) throw error("unrecognised option '" + key + "'");
        if (!e->sem) { p.given[e->long_name] = "1"; continue; }
        if (i + 1 >= argc) throw error("the required argument for option '" + key + "' is missing");
        p.given[e->long_name] = argv[++i];
    }
    return p;
}
inline void store(const parsed_options &p, variables_map &vm) {
    vm.desc = p.desc;
    for (auto &kv : p.given) vm.values[kv.first].text = kv.second;
    for (auto &e : p.desc->entries)
        if (e.sem && e.sem->has_default && !vm.values.count(e.long_name)) vm.values[e.long_name].text = e.sem->default_text;
}
inline void notify(variables_map &vm) {
    if (!vm.desc) return;
    for (auto &e : vm.desc->entries)
        if (e.sem && e.sem->is_required && !vm.values.count(e.long_name)) throw error("the option '--" + e.long_name + "' is required but missing");
}
} }
#endif
