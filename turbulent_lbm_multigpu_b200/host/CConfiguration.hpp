/*
 * CConfiguration.hpp -- run configuration, filled from the command line or from conf.xml.
 *
 * Same public fields and loadFile()/printMe() as the reference's src/CConfiguration.hpp:20-100
 * and the same schema (conf.xml:1-56, including the tag spelling `domian-length`).  The
 * reference parses with tinyxml2, a submodule that is absent from the tree (pinned 8224e42);
 * this file carries a small reader for the subset of XML the schema needs (elements, text,
 * comments, the <?xml?> declaration).  Extension: optional `physics/smagorinsky-constant`
 * (default 0 = the reference's plain BGK), so reference files keep working unchanged.
 */
#ifndef LBM_B200_HOST_CCONFIGURATION_HPP
#define LBM_B200_HOST_CCONFIGURATION_HPP

#include <cstdlib>
#include <fstream>
#include <iostream>
#include <list>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "CVector.hpp"
#include "Singleton.hpp"

namespace lbm_xml {

struct Node {
	std::string name, text;
	std::vector<Node> children;
	const Node *child(const std::string &n) const
	{
		for (size_t i = 0; i < children.size(); i++) if (children[i].name == n) return &children[i];
		return NULL;
	}
};

/* recursive-descent reader: <a> text <b>..</b> </a>, <!-- --> and <? ?> skipped, attributes ignored */
class Reader {
	const std::string &s;
	size_t p;

	void skip_misc()
	{
		for (;;) {
			while (p < s.size() && isspace((unsigned char)s[p])) p++;
			if (s.compare(p, 4, "<!--") == 0) { size_t e = s.find("-->", p); p = e == std::string::npos ? s.size() : e + 3; }
			else if (s.compare(p, 2, "<?") == 0) { size_t e = s.find("?>", p); p = e == std::string::npos ? s.size() : e + 2; }
			else return;
		}
	}

public:
	explicit Reader(const std::string &src) : s(src), p(0) {}

	bool element(Node &out)
	{
		skip_misc();
		if (p >= s.size() || s[p] != '<' || s.compare(p, 2, "</") == 0) return false;
		size_t e = s.find('>', p);
		if (e == std::string::npos) return false;
		std::string tag = s.substr(p + 1, e - p - 1);
		const bool empty = !tag.empty() && tag[tag.size() - 1] == '/';
		if (empty) tag.erase(tag.size() - 1);
		out.name = tag.substr(0, tag.find_first_of(" \t\r\n"));
		p = e + 1;
		if (empty) return true;
		for (;;) {
			size_t lt = s.find('<', p);
			if (lt == std::string::npos) return false;
			out.text += s.substr(p, lt - p);
			p = lt;
			if (s.compare(p, 4, "<!--") == 0 || s.compare(p, 2, "<?") == 0) { skip_misc(); continue; }
			if (s.compare(p, 2, "</") == 0) {
				size_t c = s.find('>', p);
				if (c == std::string::npos || s.substr(p + 2, c - p - 2).find(out.name) != 0) return false;
				p = c + 1;
				return true;
			}
			Node ch;
			if (!element(ch)) return false;
			out.children.push_back(ch);
		}
	}
};

} /* namespace lbm_xml */

template <typename T>
class CConfiguration {
public:
	/* grid */
	CVector<3, int> domain_size;
	CVector<3, int> subdomain_num;
	CVector<3, T> domain_length;
	/* physics */
	CVector<3, T> gravitation;
	T viscosity;
	CVector<4, T> drivenCavityVelocity;
	T smagorinsky_constant;            /* extension, 0 = BGK */
	/* device */
	size_t computation_kernel_count;
	int device_nr;
	/* simulation */
	bool do_visualization;
	T timestep;
	int loops;
	bool do_validate;
	std::list<int> lbm_opencl_number_of_registers_list;
	std::list<int> lbm_opencl_number_of_threads_list;
	bool debug_mode;

	/* defaults of the reference's command line (src/main.cpp:100-115) */
	CConfiguration()
		: domain_size(32, 32, 32), subdomain_num(1, 1, 1), domain_length((T)0.1, (T)0.1, (T)0.1),
		  gravitation((T)0, (T)-9.81, (T)0), viscosity((T)0.001308), drivenCavityVelocity((T)100, (T)0, (T)0, (T)1),
		  smagorinsky_constant((T)0), computation_kernel_count(128), device_nr(0), do_visualization(false),
		  timestep((T)-1.0), loops(-1), do_validate(false), debug_mode(false) {}

	explicit CConfiguration(std::string file_name) : CConfiguration() { loadFile(file_name); }

	void loadFile(std::string file_name)
	{
		std::ifstream in(file_name.c_str());
		if (!in) throw "Loading XML file failed";
		std::stringstream buf;
		buf << in.rdbuf();
		loadString(buf.str());
	}

	void loadString(const std::string &xml)
	{
		lbm_xml::Reader rd(xml);
		lbm_xml::Node root;
		if (!rd.element(root) || root.name != "lbm-configuration") throw "Loading XML file failed";
		const lbm_xml::Node &dev = need(root, "device"), &grid = need(root, "grid");
		const lbm_xml::Node &phys = need(root, "physics"), &sim = need(root, "simulation");
		computation_kernel_count = (size_t)atoi(text(dev, "kernel-count"));
		device_nr = atoi(text(dev, "device-number"));
		static const char *xyz[4] = { "x", "y", "z", "w" };
		for (int a = 0; a < 3; a++) {
			domain_size[a] = atoi(text(need(grid, "domain-size"), xyz[a]));
			subdomain_num[a] = atoi(text(need(grid, "subdomain-num"), xyz[a]));
			domain_length[a] = (T)atof(text(need(grid, "domian-length"), xyz[a]));
			gravitation[a] = (T)atof(text(need(phys, "gravitation"), xyz[a]));
		}
		viscosity = (T)atof(text(phys, "viscosity"));
		for (int a = 0; a < 4; a++) drivenCavityVelocity[a] = (T)atof(text(need(phys, "cavity-velocity"), xyz[a]));
		if (phys.child("smagorinsky-constant")) smagorinsky_constant = (T)atof(text(phys, "smagorinsky-constant"));
		loops = atoi(text(sim, "loops"));
		timestep = (T)atof(text(sim, "timestep"));
		do_visualization = atoi(text(need(sim, "visualization"), "VTK")) != 0;
		do_validate = atoi(text(sim, "validate")) != 0;
	}

	void printMe()
	{
		std::cout << "################" << std::endl << "# CONFIGURATION " << std::endl << "################" << std::endl;
		std::cout << "PHYSICS: " << std::endl;
		std::cout << "	    VISCOSITY: " << viscosity << std::endl;
		std::cout << "	  GRAVITATION: " << gravitation << std::endl;
		std::cout << "     CAVITY VEL: " << drivenCavityVelocity << std::endl;
		std::cout << "    SMAGORINSKY: " << smagorinsky_constant << std::endl;
		std::cout << "GRID: " << std::endl;
		std::cout << "	  DOMAIN_SIZE: " << domain_size << std::endl;
		std::cout << "	SUBDOMIAN_NUM: " << subdomain_num << std::endl;
		std::cout << "SIMULATION: " << std::endl;
		std::cout << "	        LOOPS: " << loops << std::endl;
		std::cout << "	     TIMESTEP: " << timestep << std::endl;
		std::cout << "	          VTK: " << do_visualization << std::endl;
		std::cout << "	     VALIDATE: " << do_validate << std::endl;
		std::cout << "DEVICE: " << std::endl;
		std::cout << "  KERNEL_COUNT: " << computation_kernel_count << std::endl;
		std::cout << "	    DEVICE_NR: " << device_nr << std::endl;
	}

private:
	static const lbm_xml::Node &need(const lbm_xml::Node &n, const char *name)
	{
		const lbm_xml::Node *c = n.child(name);
		if (!c) throw "Loading XML file failed";
		return *c;
	}
	static const char *text(const lbm_xml::Node &n, const char *name) { return need(n, name).text.c_str(); }
};

#endif
